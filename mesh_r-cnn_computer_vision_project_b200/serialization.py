"""Mesh / voxel file formats either side of the hot path: Wavefront ``.obj`` meshes and ``.npy`` /
``.mat`` / ``.binvox`` voxel grids, with the semantics of the reference ``utils/serialization.py`` (:13-41 writers,
:95-138 readers) so that files written by either implementation load identically in the other.  Host-side only.

Reference behaviour kept on purpose:
  * ``save_voxels`` stores the *mask* ``(voxels > threshold).astype(int32)`` (:13-18);
  * ``save_mesh`` appends ``.obj`` to the file name, writes ``v x y z`` then ``f i j k`` lines with 1-based indices
    (faces that already start at 1 are written unchanged, :33-35) and numpy's ``%s`` formatting of the values (:21-41);
  * ``load_mesh`` splits a polygon ``f a b c d ...`` into the *sliding* triples (a,b,c), (b,c,d), ... (:118-120 -- a strip,
    not a fan), keeps only the vertex index of ``i/j/k`` tokens, and shifts the indices to 0-based when the smallest is 1.
"""
from collections import namedtuple
from typing import Union

import numpy as np
import torch
from torch import Tensor

Mesh = namedtuple("Mesh", ["vertices", "faces"])


def _to_numpy(x) -> np.ndarray:
    return x if isinstance(x, np.ndarray) else x.detach().cpu().numpy()


def save_voxels(voxels: Union[Tensor, np.ndarray], filename: str, threshold: float = 0.5) -> None:
    np.save(filename, (_to_numpy(voxels) > threshold).astype(np.int32))


def save_mesh(vertices: Union[Tensor, np.ndarray], faces: Union[Tensor, np.ndarray], filename: str) -> None:
    vertices, faces = _to_numpy(vertices), _to_numpy(faces)
    if faces.min() == 0:
        faces = faces + 1
    rows = np.vstack((np.hstack((np.full([vertices.shape[0], 1], "v"), vertices)),
                      np.hstack((np.full([faces.shape[0], 1], "f"), faces))))
    np.savetxt(filename + ".obj", rows, fmt="%s", delimiter=" ")


def _read_binvox(fp) -> np.ndarray:
    """Run-length encoded binvox grid -> (x, y, z)-ordered 0/1 array (reference :44-92: the file stores x-z-y order and
    ``fix_coords`` transposes it)."""
    fp.readline()                                                    # '#binvox 1'
    dims = list(map(int, fp.readline().strip().split(b" ")[1:]))
    fp.readline()                                                    # translate
    fp.readline()                                                    # scale
    fp.readline()                                                    # 'data'
    raw = np.frombuffer(fp.read(), dtype=np.uint8)
    values, counts = raw[::2], raw[1::2]
    data = np.repeat(values, counts).astype(bool).reshape(dims)
    return 1 * np.transpose(data, (0, 2, 1))


def load_voxels(path: str, tensor: bool = False):
    if path.endswith(".npy"):
        vxls = np.load(path)
    elif path.endswith(".mat"):
        import scipy.io
        vxls = scipy.io.loadmat(path)["voxel"]
    else:
        assert path.endswith(".binvox")
        with open(path, "rb") as f:
            vxls = _read_binvox(f)
    return torch.from_numpy(vxls) if tensor else vxls


def load_mesh(filename: str, tensor: bool = False) -> Mesh:
    triangles, vertices = [], []
    filename = filename.replace(".binvox", ".obj")
    with open(filename) as f:
        for line in f:
            parts = line.strip(" \n").split(" ")
            if parts[0] == "f":
                idx = [int(c.split("/")[0]) for c in parts[1:]]
                triangles += [idx[i:i + 3] for i in range(len(idx) - 2)]
            elif parts[0] == "v":
                vertices.append([float(c) for c in parts[1:]])
    vertices, triangles = np.array(vertices), np.array(triangles)
    if triangles.min() == 1:
        triangles -= 1
    assert triangles.min() == 0
    if tensor:
        return Mesh(torch.from_numpy(vertices).float(), torch.from_numpy(triangles).long())
    return Mesh(vertices, triangles)

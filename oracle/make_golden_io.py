"""TEST INFRASTRUCTURE ONLY -- regenerates the mesh / voxel IO fixtures under ``tests/golden/`` (``io_*.obj``, ``io_*.npy``,
``io_expected.npz``) by running the UNMODIFIED reference ``utils/serialization.py`` (loaded by file path: its package
``__init__`` pulls in matplotlib) in the dev container.  Fixture contents are outputs of the reference, never its source.

    python -m oracle.make_golden_io         # from the repo root, dev container only (needs /root/reference)
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cubify_np, ref_import  # noqa: E402
from meshrcnn_b200 import synthetic  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("ref_serialization", os.path.join(ref_import.REF_ROOT, "utils", "serialization.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = os.path.join(ROOT, "tests", "golden")
    vox = synthetic.blob_voxels(1, 6, 3).numpy()
    v, vi, f, fi, adj = cubify_np.cubify(vox, 0.2)
    ref.save_mesh(v, f, os.path.join(g, "io_mesh_ref"))                      # -> io_mesh_ref.obj (1-based faces)
    ref.save_voxels(vox[0], os.path.join(g, "io_voxels_ref.npy"), 0.2)
    with open(os.path.join(g, "io_poly.obj"), "w") as fh:                     # polygon + i/j/k token syntax
        fh.write("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0.5 0.5 1\nf 1/1/1 2/2/2 3/3/3 4/4/4\nf 1 2 5\n")
    poly = ref.load_mesh(os.path.join(g, "io_poly.obj"))
    mesh = ref.load_mesh(os.path.join(g, "io_mesh_ref.obj"))
    np.savez_compressed(os.path.join(g, "io_expected.npz"), poly_v=poly.vertices, poly_f=poly.faces, mesh_v=mesh.vertices,
                        mesh_f=mesh.faces, verts=v, faces=f, vox=vox[0],
                        vox_mask=ref.load_voxels(os.path.join(g, "io_voxels_ref.npy")))
    print("wrote io fixtures:", poly.faces.tolist(), mesh.vertices.shape, mesh.faces.shape)


if __name__ == "__main__":
    main()

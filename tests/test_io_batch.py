"""Host-side formats either side of the hot path: ``.obj`` / ``.npy`` / ``.binvox`` IO against files
and arrays produced by the UNMODIFIED reference (``oracle/make_golden_io.py``), and the packed ``Batch`` container
(reference data/dataloader.py:11-77).  CPU only."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_obj_reader_matches_reference_outputs():
    from meshrcnn_b200.serialization import load_mesh
    g = np.load(os.path.join(GOLD, "io_expected.npz"))
    poly = load_mesh(os.path.join(GOLD, "io_poly.obj"))
    assert np.array_equal(poly.vertices, g["poly_v"]) and np.array_equal(poly.faces, g["poly_f"])
    assert poly.faces.tolist() == [[0, 1, 2], [1, 2, 3], [0, 1, 4]]          # sliding triples, not a fan (:118-120)
    mesh = load_mesh(os.path.join(GOLD, "io_mesh_ref.obj"), tensor=True)
    assert mesh.vertices.dtype == torch.float32 and mesh.faces.dtype == torch.int64
    assert np.array_equal(mesh.vertices.numpy(), g["mesh_v"].astype(np.float32)) and np.array_equal(mesh.faces.numpy(), g["mesh_f"])
    assert np.array_equal(mesh.faces.numpy(), g["faces"])                    # what Cubify emitted, 0-based again


def test_obj_writer_is_byte_identical_to_the_reference(tmp_path):
    from meshrcnn_b200.serialization import save_mesh, load_mesh
    g = np.load(os.path.join(GOLD, "io_expected.npz"))
    save_mesh(torch.from_numpy(g["verts"]), torch.from_numpy(g["faces"]), str(tmp_path / "m"))
    assert (tmp_path / "m.obj").read_bytes() == open(os.path.join(GOLD, "io_mesh_ref.obj"), "rb").read()
    save_mesh(g["verts"], g["faces"] + 1, str(tmp_path / "one_based"))      # already 1-based: written unchanged (:33-35)
    again = load_mesh(str(tmp_path / "one_based.obj"))
    assert np.array_equal(again.faces, g["faces"])


def test_voxel_files(tmp_path):
    from meshrcnn_b200.serialization import save_voxels, load_voxels
    g = np.load(os.path.join(GOLD, "io_expected.npz"))
    save_voxels(torch.from_numpy(g["vox"]), str(tmp_path / "v.npy"), 0.2)
    mine = load_voxels(str(tmp_path / "v.npy"))
    assert mine.dtype == np.int32 and np.array_equal(mine, g["vox_mask"])
    assert np.array_equal(load_voxels(os.path.join(GOLD, "io_voxels_ref.npy"), tensor=True).numpy(), g["vox_mask"])
    # binvox: run-length encoded, stored x-z-y, returned x-y-z (:44-92)
    dims = (2, 3, 4)
    data = (np.arange(24).reshape(dims) % 3 == 0)
    flat = data.reshape(-1).astype(np.uint8)
    runs, i = bytearray(), 0
    while i < flat.size:
        j = i
        while j < flat.size and flat[j] == flat[i] and j - i < 255:
            j += 1
        runs += bytes([int(flat[i]), j - i])
        i = j
    path = tmp_path / "m.binvox"
    path.write_bytes(b"#binvox 1\ndim 2 3 4\ntranslate 0 0 0\nscale 1\ndata\n" + bytes(runs))
    got = load_voxels(str(path))
    assert got.shape == (2, 4, 3) and np.array_equal(got, 1 * np.transpose(data, (0, 2, 1)))


def test_batch_container_layout_and_slicing():
    from meshrcnn_b200.batch import Batch, resample_voxels
    from meshrcnn_b200.serialization import Mesh
    g = torch.Generator().manual_seed(0)
    meshes = [Mesh(torch.rand(n, 3, generator=g), torch.randint(0, n, (f, 3), generator=g)) for n, f in ((5, 7), (9, 4), (3, 2))]
    images = torch.rand(3, 3, 8, 8, generator=g)
    voxels = [(torch.rand(4, 4, 4, generator=g) > 0.5).float() for _ in range(3)]
    targets = torch.tensor([2, 0, 1])
    b = Batch(images, voxels, 8, meshes, targets)
    assert len(b) == 3 and b.voxels.shape == (3, 8, 8, 8)
    assert torch.equal(b.voxels, resample_voxels(torch.stack(voxels), 8))
    assert torch.equal(b.voxels[:, ::2, ::2, ::2], torch.stack(voxels))       # nearest up-sampling by 2
    assert b.vertice_index == [5, 9, 3] and b.face_index == [7, 4, 2] and b.mesh_index == [1, 1, 1]
    assert b.meshes.vertices.shape == (17, 3) and b.meshes.faces.shape == (13, 3)
    assert torch.equal(b.meshes.faces[7:11], meshes[1].faces)                 # local ids, not re-offset
    sub = b[1:3]
    assert len(sub) == 2 and sub.vertice_index == [9, 3] and torch.equal(sub.meshes.vertices[9:], meshes[2].vertices)
    one = b[0]
    assert len(one) == 1 and one.face_index == [7] and torch.equal(one.backbone_targets, targets[0:1])
    down = resample_voxels(b.voxels, 4)                                        # adaptive max-pool down
    assert torch.equal(down, torch.stack(voxels))
    assert b.to(torch.float64).meshes.vertices.dtype == torch.float64
    # the losses read their ground truth through exactly these fields (loss_functions.py:54-59)
    from meshrcnn_b200.pipeline import MeshTargets
    t = MeshTargets(b.meshes.vertices, b.meshes.faces, b.vertice_index, b.face_index)
    assert t.meshes[0] is b.meshes.vertices and t.face_index == b.face_index

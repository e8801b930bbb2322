"""TEST INFRASTRUCTURE ONLY -- CPU oracle for Cubify (numpy, integer / topology work, bit-exact).

Restates the *algorithm* of the reference ``Cubify.forward`` (``meshRCNN/layers.py:403-484``) without its
conv3d / argsort / unique / Python-dict machinery: every output of the reference is a scan order over the
voxel grid or the (Z+1)x(Y+1)x(X+1) corner lattice, so the oracle computes flags and prefix sums.

Parity is pinned (``tests/test_oracle_golden.py``) against
  * the reference's own shipped golden pair ``shapenet_ex/00_voxel_obj0.npy`` ->
    ``shapenet_ex/00_mesh_stage0_obj_0.obj`` (vertices and faces, values and order), and
  * outputs of the reference itself (run in the dev container behind ``oracle/ref_import.py``) on seeded
    random / blob / edge-case grids, committed under ``tests/golden/cubify_*.npz`` by ``oracle/make_golden.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module.  The product path (``meshrcnn_b200``) never does.
"""
import numpy as np

# Neighbour whose emptiness exposes a face, (dz, dy, dx) per direction -- layers.py:357-362 (conv weights -1).
NEIGHBOUR = np.array([(-1, 0, 0), (+1, 0, 0), (0, +1, 0), (0, -1, 0), (0, 0, -1), (0, 0, +1)], dtype=np.int64)

# Quad corners c0..c3 per direction as lattice offsets in {0,1}^3 (0 = -1/2, 1 = +1/2), order (z, y, x) --
# layers.py:370-400.  NOTE dirs 2/3: the quad sits on the side *opposite* the empty neighbour (reference quirk).
CORNERS = np.array([
    [(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1)],   # 0 back   (z-1 empty) plane z-1/2
    [(1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)],   # 1 front  (z+1 empty) plane z+1/2
    [(1, 0, 0), (1, 0, 1), (0, 0, 0), (0, 0, 1)],   # 2 top    (y+1 empty) plane y-1/2
    [(0, 1, 0), (0, 1, 1), (1, 1, 0), (1, 1, 1)],   # 3 bottom (y-1 empty) plane y+1/2
    [(1, 0, 0), (0, 0, 0), (1, 1, 0), (0, 1, 0)],   # 4 left   (x-1 empty) plane x-1/2
    [(0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)],   # 5 right  (x+1 empty) plane x+1/2
], dtype=np.int64)

# Two (overlapping) triangles per quad -- layers.py:441-443.
TRIANGLES = ((0, 1, 2), (0, 2, 3))


class EmptyGrid(ValueError):
    pass


def occupancy(probs: np.ndarray, threshold: float) -> np.ndarray:
    """layers.py:405 -- strict '>' evaluated in fp32 (the python float threshold is cast to the tensor dtype)."""
    p = np.asarray(probs)
    if p.dtype != np.float32:
        p = p.astype(np.float32)
    return p > np.float32(threshold)


def face_flags(occ: np.ndarray) -> np.ndarray:
    """(B,6,Z,Y,X) bool: occupied voxel whose neighbour in direction d is empty / outside (zero padding,
    layers.py:411-415: conv output == 1  <=>  centre 1 and neighbour 0)."""
    B, Z, Y, X = occ.shape
    pad = np.zeros((B, Z + 2, Y + 2, X + 2), dtype=bool)
    pad[:, 1:-1, 1:-1, 1:-1] = occ
    out = np.empty((B, 6, Z, Y, X), dtype=bool)
    for d, (dz, dy, dx) in enumerate(NEIGHBOUR):
        nb = pad[:, 1 + dz:1 + dz + Z, 1 + dy:1 + dy + Y, 1 + dx:1 + dx + X]
        out[:, d] = occ & ~nb
    return out


def cubify(probs: np.ndarray, threshold: float = 0.5):
    """Returns (verts float32 [SV,3], v_index list, faces int64 [SF,3] per-mesh-local, f_index list,
    adj int64 [2,E] global) exactly like ``Cubify(threshold).forward`` (layers.py:484).

    Orders (probed): vertices by (b,z,y,x); faces by (b,dir,z,y,x), two per quad;
    adjacency by (row,col)."""
    occ = occupancy(probs, threshold)
    B, Z, Y, X = occ.shape
    flags = face_flags(occ)
    b_i, d_i, z_i, y_i, x_i = np.nonzero(flags)        # C order == (b, dir, z, y, x)
    nquads = b_i.shape[0]
    if nquads == 0:
        raise EmptyGrid("empty grid")                   # layers.py:434-435

    LZ, LY, LX = Z + 1, Y + 1, X + 1
    corner = CORNERS[d_i]                               # (N,4,3)
    cz = z_i[:, None] + corner[:, :, 0]
    cy = y_i[:, None] + corner[:, :, 1]
    cx = x_i[:, None] + corner[:, :, 2]
    lin = ((b_i[:, None] * LZ + cz) * LY + cy) * LX + cx  # (N,4) lattice linear ids

    used = np.zeros(B * LZ * LY * LX, dtype=bool)
    used[lin.ravel()] = True
    rank = np.cumsum(used, dtype=np.int64) - 1            # vertex id = lexicographic rank (layers.py:447)
    ids = np.nonzero(used)[0]
    vb = ids // (LZ * LY * LX)
    rem = ids % (LZ * LY * LX)
    vz = rem // (LY * LX)
    vy = (rem // LX) % LY
    vx = rem % LX
    fz = vz.astype(np.float32) - np.float32(0.5)
    fy = vy.astype(np.float32) - np.float32(0.5)
    fx = vx.astype(np.float32) - np.float32(0.5)
    verts = np.stack([fz, fx, -fy], axis=1).astype(np.float32)   # rotation(90): (z,y,x) -> (z,x,-y), :465-467

    v_counts = np.bincount(vb)                            # truncated after the last non-empty mesh (:448)
    f_counts = np.bincount(b_i) * 2                       # (:445)
    v_index = [int(c) for c in v_counts]
    f_index = [int(c) for c in f_counts]

    g = rank[lin]                                         # (N,4) global vertex ids
    tri = np.stack([g[:, TRIANGLES[0]], g[:, TRIANGLES[1]]], axis=1).reshape(-1, 3)   # (2N,3)
    v_off = np.concatenate([[0], np.cumsum(v_counts)])[:-1]
    v_off_full = np.zeros(B, dtype=np.int64)
    v_off_full[:len(v_off)] = v_off
    faces = tri - v_off_full[np.repeat(b_i, 2)][:, None]  # per-mesh local ids (:481-483)

    # symmetric adjacency from the three edges of every triangle (:469-478), sorted by (row, col)
    e_r = np.concatenate([tri[:, 0], tri[:, 1], tri[:, 0]])
    e_c = np.concatenate([tri[:, 1], tri[:, 2], tri[:, 2]])
    rows = np.concatenate([e_r, e_c])
    cols = np.concatenate([e_c, e_r])
    nv = ids.shape[0]
    key = np.unique(rows * nv + cols)
    adj = np.stack([key // nv, key % nv]).astype(np.int64)
    return verts, v_index, faces.astype(np.int64), f_index, adj

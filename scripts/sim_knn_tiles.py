"""CPU simulation (numpy) of candidate-tile pruning policies for the k-NN kernel: how many candidate points does a
group of G consecutive queries have to visit when both clouds are ordered along a space-filling curve and candidate
tiles of T points are skipped by their bounding box?  Design evidence for csrc/chamfer.cu (no GPU needed).

    python scripts/sim_knn_tiles.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cubify_np, mesh_ops                      # noqa: E402
from meshrcnn_b200 import synthetic                         # noqa: E402


def cloud(seed, th, n=10000, V=24, jitter=0.0):
    vox = synthetic.blob_voxels(1, V, seed).numpy()
    v, vi, f, fi, _ = cubify_np.cubify(vox, th)
    v = torch.from_numpy(np.asarray(v, dtype=np.float32))
    f = torch.from_numpy(np.asarray(f, dtype=np.int64))
    if jitter:
        v = v + jitter * torch.rand(v.shape, generator=torch.Generator().manual_seed(seed))
    g = torch.Generator().manual_seed(seed + 5)
    u = torch.rand(n, generator=g)
    fidx = mesh_ops.face_cdf_draw(v, f, u)
    pts = mesh_ops.sample_with(v, f, fidx, torch.rand(n, generator=g), torch.rand(n, generator=g))
    return mesh_ops.normalize_cloud(pts).numpy().astype(np.float32)


def morton(pts, bits=10):
    lo, hi = pts.min(0), pts.max(0)
    q = np.minimum(((pts - lo) / (hi - lo + 1e-9) * (1 << bits)).astype(np.int64), (1 << bits) - 1)
    code = np.zeros(len(pts), dtype=np.int64)
    for b in range(bits):
        for d in range(3):
            code |= ((q[:, d] >> b) & 1) << (3 * b + d)
    return code


def order(pts, kind):
    if kind == "x":
        return np.argsort(pts[:, 0], kind="stable")
    if kind == "morton":
        return np.argsort(morton(pts), kind="stable")
    raise ValueError(kind)


def simulate(a, b, kind, G, T, k):
    a = a[order(a, kind)]
    b = b[order(b, kind)]
    nt = (len(b) + T - 1) // T
    tlo = np.stack([b[t * T:(t + 1) * T].min(0) for t in range(nt)])
    thi = np.stack([b[t * T:(t + 1) * T].max(0) for t in range(nt)])
    visited = 0
    tiles = 0
    for g0 in range(0, len(a), G):
        qs = a[g0:g0 + G]
        qlo, qhi = qs.min(0), qs.max(0)
        if kind == "x":
            gap = np.maximum(0, np.maximum(tlo[:, 0] - qhi[0], qlo[0] - thi[:, 0])) ** 2
        else:
            gap = (np.maximum(0, np.maximum(tlo - qhi, qlo - thi)) ** 2).sum(1)
        idx = np.argsort(gap, kind="stable")
        best = np.full((len(qs), k), np.inf)
        for t in idx:
            if gap[t] > best[:, -1].max():
                break
            c = b[t * T:(t + 1) * T]
            d = ((qs[:, None, :] - c[None]) ** 2).sum(-1)
            best = np.sort(np.concatenate([best, d], 1), 1)[:, :k]
            visited += len(c) * len(qs)
            tiles += 1
    return visited / (len(a) * len(b)), tiles / ((len(a) + G - 1) // G)


if __name__ == "__main__":
    p = cloud(0, 0.2, jitter=1.0)      # "prediction": cubified blob with O(1)-voxel offsets, as a random head gives
    q = cloud(1000, 0.5)               # ground truth
    for k in (1, 10):
        for kind, G, T in (("x", 64, 256), ("morton", 64, 256), ("morton", 64, 64), ("morton", 32, 64),
                           ("morton", 32, 32), ("morton", 32, 128)):
            fr, tl = simulate(p, q, kind, G, T, k)
            print("k=%2d  order=%-6s  queries/group=%3d  tile=%3d  visited %.4f of pairs, %.1f tiles/group" %
                  (k, kind, G, T, fr, tl))

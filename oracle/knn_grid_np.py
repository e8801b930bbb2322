"""TEST INFRASTRUCTURE ONLY -- numpy (float32) restatement of the exact cell-grid k-NN search of ``csrc/chamfer.cu``
(``k_grid_build`` + ``k_nn_grid``): counting sort of the candidate cloud into G^3 cells, per-query cell box for radius r,
"final iff K keys were found and the K-th distance <= r^2, else r := that distance (or 2 r) and restart with the old K-th
distance as a filter".  It replaces nothing in the reference (which materialises the B x P x Q matrix and calls ``min`` /
``topk``, meshRCNN/loss_functions.py:93-102,141); it exists so that the *exactness argument* of the search -- a monotone
``cell_of`` shared by points and box corners, a box radius 1e-5 above r -- is pinned on the CPU against brute force,
independently of the CUDA kernel (which the GPU tests compare bit for bit with the tiled scan and with the fp64 oracle).
"""
import numpy as np

GMAX = 32
C0 = 4.0
f32 = np.float32


def grid_cells(n: int) -> int:
    return int(min(GMAX, max(1, np.ceil(np.sqrt(n / 12.0)))))


def cell_of(x, lo, inv, G):
    """min(G - 1, max(0, floor(fl(fl(x - lo) * inv))))  -- every operation in fp32, like the kernel's __fsub_rn / __fmul_rn."""
    t = (np.asarray(x, dtype=f32) - f32(lo)).astype(f32) * f32(inv)
    with np.errstate(invalid="ignore"):
        c = np.floor(t.astype(f32))
    c = np.where(np.isnan(c), 0, c)
    return np.clip(c, 0, G - 1).astype(np.int64)


def build(points: np.ndarray, k: int):
    """-> dict(lo, inv, G, r0, start (G^3 + 1 offsets, x fastest), order (sorted -> original index), pts (sorted))."""
    pts = np.asarray(points, dtype=f32)
    n = len(pts)
    G = grid_cells(n)
    lo, hi = pts.min(0), pts.max(0)
    ext = (hi - lo).astype(f32)
    inv = np.where(ext > 0, f32(G) / np.where(ext > 0, ext, 1), 0).astype(f32)
    c = np.stack([cell_of(pts[:, d], lo[d], inv[d], G) for d in range(3)], 1)
    cid = (c[:, 2] * G + c[:, 1]) * G + c[:, 0]
    order = np.argsort(cid, kind="stable")
    start = np.zeros(G ** 3 + 1, dtype=np.int64)
    np.add.at(start, cid + 1, 1)
    start = np.cumsum(start)
    occupied = int((np.diff(start) > 0).sum())
    hs = [ext[d] / f32(G) for d in range(3) if inv[d] > 0]
    hmean = f32(np.mean(hs)) if hs else f32(0)
    r0 = max(f32(0.5) * hmean * f32(np.sqrt(C0 * max(k, 1))) * f32(np.sqrt(max(occupied, 1) / max(n, 1))), f32(1e-20))
    return dict(lo=lo, inv=inv, G=G, r0=f32(r0), start=start, order=order, pts=pts[order])


def knn(queries: np.ndarray, grid: dict, K: int):
    """(dist K, idx K) per query, sorted by (distance, original index); idx = -1 / dist = inf where fewer than K points exist."""
    q = np.asarray(queries, dtype=f32)
    G, lo, inv, start, pts, order = grid["G"], grid["lo"], grid["inv"], grid["start"], grid["pts"], grid["order"]
    out_d = np.full((len(q), K), np.inf, dtype=f32)
    out_i = np.full((len(q), K), -1, dtype=np.int64)
    visited = np.zeros(len(q), dtype=np.int64)
    for n_, p in enumerate(q):
        r = grid["r0"]
        thr = (f32(np.inf), np.iinfo(np.int64).max)
        for _ in range(512):
            rr = f32(r * f32(1e-5) + r) + f32(1e-30)
            c0 = [int(cell_of(p[d] - rr, lo[d], inv[d], G)) for d in range(3)]
            c1 = [int(cell_of(p[d] + rr, lo[d], inv[d], G)) for d in range(3)]
            cand = []
            for cz in range(c0[2], c1[2] + 1):
                for cy in range(c0[1], c1[1] + 1):
                    row = (cz * G + cy) * G
                    cand.append(np.arange(start[row + c0[0]], start[row + c1[0] + 1]))
            cand = np.concatenate(cand) if cand else np.zeros(0, dtype=np.int64)
            visited[n_] += len(cand)
            diff = (p[None, :] - pts[cand]).astype(f32)
            d = (diff[:, 2] * diff[:, 2] + (diff[:, 1] * diff[:, 1] + diff[:, 0] * diff[:, 0]).astype(f32)).astype(f32)
            idx = order[cand]
            keep = (d < thr[0]) | ((d == thr[0]) & (idx < thr[1]))
            d, idx = d[keep], idx[keep]
            sel = np.lexsort((idx, d))[:K]
            full = len(sel) == K
            if full and d[sel[-1]] <= f32(r * r):
                break
            whole = all(inv[d_] == 0 or (c0[d_] == 0 and c1[d_] == G - 1) for d_ in range(3))
            if whole:
                break
            if full:
                kd = d[sel[-1]]
                r = f32(np.sqrt(kd)) * f32(1e-6) + f32(np.sqrt(kd))
                thr = (kd, np.iinfo(np.int64).max)        # restart: every key with d <= kd is admitted again
            else:
                r = f32(2) * r
        out_d[n_, :len(sel)] = d[sel]
        out_i[n_, :len(sel)] = idx[sel]
    return out_d, out_i, visited

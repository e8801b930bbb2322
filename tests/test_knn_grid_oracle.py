"""The exactness argument of the cell-grid k-NN search (csrc/chamfer.cu: k_grid_build + k_nn_grid), pinned on the CPU:
its numpy restatement (oracle/knn_grid_np.py) must return exactly the (distance, index)-ordered K nearest neighbours of a
brute-force scan -- on surfaces, lattices full of ties, degenerate axes, far-apart clouds, tiny clouds."""
import numpy as np
import pytest

from oracle import knn_grid_np as KG


def clouds(kind, P, Q, seed):
    g = np.random.default_rng(seed)
    if kind == "cube":
        return g.random((P, 3), dtype=np.float32) - 0.5, g.random((Q, 3), dtype=np.float32) - 0.5
    if kind == "sphere":
        a, b = g.standard_normal((P, 3)).astype(np.float32), g.standard_normal((Q, 3)).astype(np.float32)
        return a / np.linalg.norm(a, axis=1, keepdims=True), 0.7 * b / np.linalg.norm(b, axis=1, keepdims=True) * np.float32([1, 0.6, 1.4])
    if kind == "far":
        return g.random((P, 3), dtype=np.float32) + 5, g.random((Q, 3), dtype=np.float32) * 0.1
    if kind == "plane":
        a = np.stack([g.integers(0, 12, P), g.integers(0, 12, P), np.zeros(P)], 1).astype(np.float32)
        b = np.stack([g.integers(0, 12, Q), g.integers(0, 12, Q), np.zeros(Q)], 1).astype(np.float32)
        return a, b
    if kind == "point":
        return g.random((P, 3), dtype=np.float32), np.full((Q, 3), 0.25, dtype=np.float32)
    raise ValueError(kind)


@pytest.mark.parametrize("kind,P,Q,K", [("cube", 300, 2500, 10), ("sphere", 200, 6000, 10), ("far", 50, 700, 4),
                                        ("plane", 150, 1100, 10), ("point", 40, 200, 16), ("cube", 17, 16, 16),
                                        ("sphere", 100, 9000, 1), ("cube", 5, 7, 10)])
def test_grid_search_equals_brute_force(kind, P, Q, K):
    a, b = clouds(kind, P, Q, P + Q + K)
    grid = KG.build(b, K)
    d, i, visited = KG.knn(a, grid, K)
    diff = a[:, None, :] - b[None]
    full = (diff[..., 2] * diff[..., 2] + (diff[..., 1] * diff[..., 1] + diff[..., 0] * diff[..., 0])).astype(np.float32)
    for n in range(P):
        order = np.lexsort((np.arange(Q), full[n]))[:K]
        assert np.array_equal(i[n, :len(order)], order), (kind, n)
        assert np.array_equal(d[n, :len(order)], full[n, order])
        assert (i[n, len(order):] == -1).all()
    if kind == "sphere" and Q >= 6000 and K == 10:
        assert visited.mean() < 0.1 * Q                    # the point of the exercise: a few % of the candidates are visited


def test_cell_of_is_monotone_and_shared_by_points_and_box_corners():
    g = np.random.default_rng(0)
    x = np.sort(g.random(20000, dtype=np.float32) * 3 - 1)
    for G, lo, inv in ((29, np.float32(-0.37), np.float32(11.3)), (4, np.float32(0.0), np.float32(1.7)), (32, np.float32(5), np.float32(0))):
        c = KG.cell_of(x, lo, inv, G)
        assert (np.diff(c) >= 0).all() and c.min() >= 0 and c.max() <= G - 1

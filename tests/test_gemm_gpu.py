"""tcgen05 (3xTF32) projection kernel vs fp64 matmul: fp32-level accuracy (rtol 1e-4 of the north star with margin)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(M, K, N, lda_pad=0, ldc_pad=0, split=None, transposed_src=False, seed=0):
    from meshrcnn_b200 import _lib, functional as F_
    g = torch.Generator().manual_seed(seed)
    a_full = torch.randn(M, K + lda_pad, generator=g)
    a = a_full[:, :K]
    if split is None:
        w = torch.randn(N, K, generator=g) if transposed_src else torch.randn(K, N, generator=g)
        wl = w.t() if transposed_src else w
        wd = w.cuda()
        sk, sn = (1, K) if transposed_src else (N, 1)
        img = F_.tc_pack(wd, None, sk, sn, 0, 0, K, N)
    elif split == "n":          # [W0 | W1], each K x N/2
        w0, w1 = torch.randn(K, N // 2, generator=g), torch.randn(K, N // 2, generator=g)
        wl = torch.cat([w0, w1], 1)
        w0d, w1d = w0.cuda(), w1.cuda()
        img = F_.tc_pack(w0d, w1d, N // 2, 1, 1, N // 2, K, N)
    else:                       # [W0 | W1]^T : logical (k, n) = Wcat[n][k], W* are N x K/2
        w0, w1 = torch.randn(N, K // 2, generator=g), torch.randn(N, K // 2, generator=g)
        wl = torch.cat([w0, w1], 1).t()
        w0d, w1d = w0.cuda(), w1.cuda()
        img = F_.tc_pack(w0d, w1d, 1, K // 2, 2, K // 2, K, N)
    ad = a_full.cuda()
    c_all = torch.full((M + 40, N + ldc_pad), -7.0, device="cuda")     # 40 guard rows behind the matrix
    c = c_all[:M]
    F_.tc_gemm(_lib.ptr(ad), K + lda_pad, M, K, img, N, _lib.ptr(c), N + ldc_pad)
    torch.cuda.synchronize()
    assert bool((c_all[M:] == -7.0).all())                 # the (TMA) stores of the last, partial row tile are clipped at M
    want = a.double() @ wl.double()
    got = c[:, :N].cpu().double()
    scale = float(want.abs().max())
    err = float((got - want).abs().max())
    assert err <= 1e-5 * scale, (M, K, N, err, scale)
    rel = float((got - want).norm() / want.norm())
    # 3xTF32: ~6e-7 with a separate accumulator for the cross terms, up to ~2e-6 when short reductions (K <= 512) chain all
    # three products into one TMEM accumulator (256-column tiles, so that the accumulator can be double buffered); IEEE fp32: ~3e-7
    assert rel <= 4e-6, (M, K, N, rel)
    if ldc_pad:
        assert bool((c[:, N:] == -7.0).all())          # no write outside the N columns


@pytest.mark.parametrize("M,K,N", [(128, 32, 16), (128, 32, 256), (1000, 131, 256), (333, 259, 256), (4097, 387, 256),
                                   (50353, 131, 256), (257, 128, 128), (100, 3840, 128), (5, 8, 16)])
def test_tc_gemm_shapes(lib, M, K, N):
    _check(M, K, N)


def test_tc_gemm_strides_and_padding(lib):
    _check(300, 131, 128, lda_pad=5, ldc_pad=3)
    _check(300, 256, 131, ldc_pad=1)                   # N not a multiple of 16 (backward of a 131-wide layer)
    _check(700, 256, 387)                              # two N tiles of 208
    _check(700, 256, 259, transposed_src=True)


def test_tc_gemm_split_sources(lib):
    _check(513, 131, 256, split="n")                   # GraphConv forward  x @ [W0 | W1]
    _check(513, 256, 131, split="k")                   # GraphConv backward gy @ [W0 | W1]^T
    _check(513, 256, 387, split="k")


@pytest.mark.parametrize("V,Kin,N,split", [(1000, 128, 256, 128), (4097, 131, 256, 128), (50353, 387, 256, 128),
                                           (333, 259, 128, 128), (70, 16, 32, 32), (5000, 3840, 128, 128)])
def test_tc_wgrad(lib, V, Kin, N, split):
    """C0 | C1 = X^T @ G on the tensor cores (MN-major operands, split over CTAs, fp32 reductions)."""
    from meshrcnn_b200 import _lib
    g = torch.Generator().manual_seed(V + Kin)
    x, gy = torch.randn(V, Kin, generator=g), torch.randn(V, N, generator=g)
    xd, gd = x.cuda(), gy.cuda()
    c0 = torch.zeros(Kin, split, device="cuda")
    c1 = torch.zeros(Kin, N - split, device="cuda") if split < N else None
    _lib.call("mrb_gemm_tc_wgrad", _lib.ptr(xd), Kin, _lib.ptr(gd), N, V, Kin, N, _lib.ptr(c0), _lib.ptr(c1), split, split)
    torch.cuda.synchronize()
    want = x.double().t() @ gy.double()
    got = torch.cat([c0, c1], 1) if c1 is not None else c0
    err = float((got.cpu().double() - want).abs().max())
    assert err <= 1e-5 * float(want.abs().max()), (V, Kin, N, err, float(want.abs().max()))


def test_tc_gemm_accumulate_and_wgrad_split(lib):
    """mrb_gemm_tc_acc (C += A B: a product with [x_a | x_b] as two calls) and mrb_gemm_tc_wgrad_split (the 3 position
    columns from their own matrix, written to their own rows of dW0 | dW1)."""
    from meshrcnn_b200 import _lib, functional as F_
    g = torch.Generator().manual_seed(11)
    M, Ka, Kb, N = 3000, 128, 64, 256
    xa, xb = torch.randn(M, Ka, generator=g), torch.randn(M, Kb, generator=g)
    w = torch.randn(Ka + Kb, N, generator=g)
    wd, xad, xbd = w.cuda(), xa.cuda(), xb.cuda()
    c = torch.empty(M, N, device="cuda")
    lib_ = _lib.load()
    for i, (x, r, K) in enumerate(((xad, 0, Ka), (xbd, Ka, Kb))):
        img = torch.empty(lib_.mrb_gemm_tc_image_bytes(K, N), dtype=torch.uint8, device="cuda")
        _lib.call("mrb_gemm_tc_pack", wd.data_ptr() + 4 * r * N, None, N, 1, 0, 0, K, N, _lib.ptr(img))
        _lib.call("mrb_gemm_tc_acc", _lib.ptr(x), K, M, K, _lib.ptr(img), N, _lib.ptr(c), N, i)
    want = torch.cat([xa, xb], 1).double() @ w.double()
    assert float((c.cpu().double() - want).norm() / want.norm()) <= 2e-6
    # wgrad: dW rows [3, 131) from x (128 wide), rows [0, 3) from pos -- the layout of a [pos | x] layer
    V, D = 4097, 128
    x, pos, gy = torch.randn(V, 128, generator=g), torch.randn(V, 3, generator=g), torch.randn(V, 2 * D, generator=g)
    xd, pd, gd = x.cuda(), pos.cuda(), gy.cuda()
    gw = torch.zeros(2, 131, D, device="cuda")
    g0, g1 = gw.data_ptr(), gw.data_ptr() + 4 * 131 * D
    _lib.call("mrb_gemm_tc_wgrad_split", _lib.ptr(xd), 128, _lib.ptr(gd), 2 * D, V, 128, 2 * D, g0 + 4 * 3 * D, g1 + 4 * 3 * D, D, D,
              _lib.ptr(pd), 3, 3, g0, g1)
    want = torch.cat([pos, x], 1).double().t() @ gy.double()
    got = torch.cat([gw[0], gw[1]], 1).cpu().double()
    assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max())


@pytest.mark.parametrize("ta,tb,M,N,K", [(0, 1, 5000, 3, 131), (0, 0, 5000, 131, 3), (1, 0, 3, 131, 5000), (0, 1, 77, 8, 40),
                                         (1, 0, 8, 300, 999), (0, 0, 100, 50, 70), (1, 1, 33, 65, 129), (1, 0, 131, 128, 4000),
                                         (0, 0, 5000, 3, 128), (0, 1, 5000, 128, 3), (1, 0, 128, 3, 5000), (1, 0, 200, 6, 3000),
                                         (1, 0, 3, 3, 50353), (1, 0, 4, 1, 777), (1, 0, 3, 128, 50353)])
def test_sgemm_simt_all_paths(lib, ta, tb, M, N, K):
    """Exact-fp32 CUDA-core GEMM incl. the skinny special cases of the 3-wide heads, with beta accumulation."""
    from meshrcnn_b200 import _lib
    g = torch.Generator().manual_seed(M * 7 + N)
    a = torch.randn((K, M) if ta else (M, K), generator=g)
    b = torch.randn((N, K) if tb else (K, N), generator=g)
    c0 = torch.randn(M, N, generator=g)
    for beta in (0.0, 1.0):
        c = c0.clone().cuda()
        ad, bd = a.cuda(), b.cuda()
        _lib.call("mrb_sgemm", ta, tb, M, N, K, _lib.ptr(ad), a.shape[1], _lib.ptr(bd), b.shape[1], beta, _lib.ptr(c), N)
        want = (a.t() if ta else a).double() @ (b.t() if tb else b).double() + beta * c0.double()
        err = float((c.cpu().double() - want).abs().max())
        assert err <= 2e-6 * float(want.abs().max()) * max(1.0, (K / 100) ** 0.5), (ta, tb, M, N, K, beta, err)

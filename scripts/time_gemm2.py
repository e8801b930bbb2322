"""Micro-benchmark of the tcgen05 kernels at the split-input GraphConv shapes (K = 128 blocks, no padded chunk):
forward x @ [W0|W1] (K=128 -> N=256), input gradient (K=256 -> N=128), texel projection (M = 4608), weight gradient."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshrcnn_b200 import functional as F_, _lib
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(run, n=10):
    for _ in range(3): run()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n
for M in (50353, 205947):
    for (K, N) in [(128, 256), (256, 128), (256, 256)]:
        a = torch.randn(M, K, device="cuda"); w = torch.randn(K, N, device="cuda"); c = torch.empty(M, N, device="cuda")
        img = F_.tc_pack(w, None, N, 1, 0, 0, K, N)
        ms = timeit(lambda: F_.tc_gemm(_lib.ptr(a), K, M, K, img, N, _lib.ptr(c), N))
        ref = a.double() @ w.double()
        err = float((c.double() - ref).norm() / ref.norm())
        print("gemm  M=%6d K=%d N=%d  %6.1f us  %5.0f GB/s (A+C)  %5.1f TFLOP/s TF32 issued  rel.err %.1e" %
              (M, K, N, ms * 1e3, 4 * M * (K + N) / ms / 1e6, 6 * M * K * N / ms / 1e9, err))
    x = torch.randn(M, 128, device="cuda"); pos = torch.randn(M, 3, device="cuda"); gy = torch.randn(M, 256, device="cuda")
    gw = torch.zeros(2, 131, 128, device="cuda")
    g0, g1 = gw.data_ptr(), gw.data_ptr() + 4 * 131 * 128
    ms = timeit(lambda: _lib.call("mrb_gemm_tc_wgrad_split", _lib.ptr(x), 128, _lib.ptr(gy), 256, M, 128, 256, g0 + 4 * 3 * 128, g1 + 4 * 3 * 128,
                                  128, 128, _lib.ptr(pos), 3, 3, g0, g1))
    print("wgrad M=%6d Kin=128(+3) N=256  %6.1f us  %5.0f GB/s (X+G)  %5.1f TFLOP/s TF32 issued" %
          (M, ms * 1e3, 4 * M * (128 + 256) / ms / 1e6, 6 * M * 128 * 256 / ms / 1e9))

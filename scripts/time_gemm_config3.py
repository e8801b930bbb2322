"""tcgen05 projection / weight-gradient kernels at the ShapeNet bottleneck shapes (M = 206k vertices, 3840 <-> 128)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshrcnn_b200 import functional as F_, _lib
M = int(sys.argv[1]) if len(sys.argv) > 1 else 205947
def timeit(run, n=5):
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (K, N) in [(3840, 128), (128, 3840), (128, 256), (128, 128)]:
    a = torch.randn(M, K, device="cuda"); w = torch.randn(K, N, device="cuda"); c = torch.empty(M, N, device="cuda")
    img = F_.tc_pack(w, None, N, 1, 0, 0, K, N)
    ms = timeit(lambda: F_.tc_gemm(_lib.ptr(a), K, M, K, img, N, _lib.ptr(c), N))
    print("gemm  M=%d K=%d N=%d  %.1f us  %.1f GB/s (A+C)  %.1f TFLOP/s TF32 issued" % (M, K, N, ms * 1e3, 4 * M * (K + N) / ms / 1e6, 6 * M * K * N / ms / 1e9))
    del a, c
for (Kin, N) in [(3840, 128), (128, 128), (128, 256)]:
    x = torch.randn(M, Kin, device="cuda"); gy = torch.randn(M, N, device="cuda"); gw = torch.zeros(Kin, N, device="cuda")
    ms = timeit(lambda: _lib.call("mrb_gemm_tc_wgrad", _lib.ptr(x), Kin, _lib.ptr(gy), N, M, Kin, N, _lib.ptr(gw), None, N, N))
    print("wgrad V=%d Kin=%d N=%d  %.1f us  %.1f GB/s (X+G)  %.1f TFLOP/s TF32 issued" % (M, Kin, N, ms * 1e3, 4 * M * (Kin + N) / ms / 1e6, 6 * M * Kin * N / ms / 1e9))
    del x, gy

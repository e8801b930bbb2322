"""Micro-benchmark of the tcgen05 projection kernel at the Pix3D-head shapes (M = 50k vertices)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshrcnn_b200 import functional as F_, _lib
M = 50353
for (K, N) in [(131, 256), (259, 256), (387, 256), (256, 131), (256, 387)]:
    a = torch.randn(M, K, device="cuda"); w = torch.randn(K, N, device="cuda"); c = torch.empty(M, N, device="cuda")
    img = F_.tc_pack(w, None, N, 1, 0, 0, K, N)
    run = lambda: F_.tc_gemm(_lib.ptr(a), K, M, K, img, N, _lib.ptr(c), N)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    byts = 4 * M * (K + N)
    print("M=%d K=%d N=%d  %.1f us  %.1f GB/s (A+C)  %.1f TFLOP/s (fp32-equivalent)" % (M, K, N, ms * 1e3, byts / ms / 1e6, 2 * M * K * N / ms / 1e9))

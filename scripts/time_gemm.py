"""Micro-benchmark of the tcgen05 projection kernel at the Pix3D-head shapes (M = 50k vertices by default), with the
natural leading dimensions (lda = K, ldc = N) and with rows padded to 16 bytes (the layout the stage code produces)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshrcnn_b200 import functional as F_, _lib
M = int(sys.argv[1]) if len(sys.argv) > 1 else 50353
r4 = lambda v: (v + 3) // 4 * 4
for (K, N) in [(131, 256), (259, 256), (387, 256), (256, 131), (256, 387)]:
    for pad in (False, True):
        lda, ldc = (r4(K), r4(N)) if pad else (K, N)
        if pad and lda == K and ldc == N:
            continue
        a = torch.randn(M, lda, device="cuda"); w = torch.randn(K, N, device="cuda"); c = torch.empty(M, ldc, device="cuda")
        img = F_.tc_pack(w, None, N, 1, 0, 0, K, N)
        run = lambda: F_.tc_gemm(_lib.ptr(a), lda, M, K, img, N, _lib.ptr(c), ldc)
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        byts = 4 * M * (K + N)
        ref = a[:, :K].double() @ w.double()
        err = float((c[:, :N].double() - ref).norm() / ref.norm())
        print("M=%d K=%d N=%d lda=%d ldc=%d  %.1f us  %.1f GB/s (A+C)  %.1f TFLOP/s (fp32-equivalent)  rel.err %.1e" %
              (M, K, N, lda, ldc, ms * 1e3, byts / ms / 1e6, 2 * M * K * N / ms / 1e9, err))

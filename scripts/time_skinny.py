"""Micro-benchmark of mrb_sgemm's skinny TN path at the head weight-gradient shapes: C[3 x N] = gpre^T [3 x V] * x [V x N]."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshrcnn_b200 import _lib
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(run, n=10):
    for _ in range(3): run()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n
for V in (50353, 205947):
    for N in (128, 3, 256):
        a = torch.randn(V, 3, device="cuda"); b = torch.randn(V, N, device="cuda"); c = torch.zeros(3, N, device="cuda")
        run = lambda: _lib.call("mrb_sgemm", 1, 0, 3, N, V, _lib.ptr(a), 3, _lib.ptr(b), N, 0.0, _lib.ptr(c), N)
        ms = timeit(run)
        ref = a.double().t() @ b.double()
        err = float((c.double() - ref).norm() / ref.norm())
        print("skinny_tn V=%6d M=3 N=%3d  %6.1f us  %5.0f GB/s  rel.err %.1e" % (V, N, ms * 1e3, 4 * V * (3 + N) / ms / 1e6, err))

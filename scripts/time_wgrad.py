"""Micro-benchmark of the tcgen05 weight-gradient kernel at the Pix3D-head shapes."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshrcnn_b200 import _lib
V = int(sys.argv[1]) if len(sys.argv) > 1 else 50353
for Kin in (131, 259, 387, 128):
    x = torch.randn(V, Kin, device="cuda"); gy = torch.randn(V, 256, device="cuda"); gw = torch.zeros(2, Kin, 128, device="cuda")
    def run():
        _lib.call("mrb_gemm_tc_wgrad", _lib.ptr(x), Kin, _lib.ptr(gy), 256, V, Kin, 256, _lib.ptr(gw), _lib.ptr(gw) + 4 * Kin * 128, 128, 128)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("V=%d Kin=%d N=256  %.1f us  %.1f GB/s (X+G)  %.1f TFLOP/s (fp32-equivalent)" %
          (V, Kin, ms * 1e3, 4 * V * (Kin + 256) / ms / 1e6, 2 * V * Kin * 256 / ms / 1e9))

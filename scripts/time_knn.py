"""Micro-benchmark of mrb_knn_fwd at the chamfer-sweep size (BASELINE config 5: B=32, P=Q=10k)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshrcnn_b200 import functional as F_, _lib

B, P, k = 32, 10000, int(sys.argv[1]) if len(sys.argv) > 1 else 10
g = torch.Generator().manual_seed(0)
def cloud():
    x = torch.randn(B, P, 3, generator=g)
    return (x / x.norm(dim=2, keepdim=True) * torch.rand(B, P, 1, generator=g) ** 0.5).cuda()
p, q = cloud(), cloud()
dp = torch.empty(B, P, device="cuda"); ip = torch.empty(B, P, dtype=torch.int32, device="cuda")
kp = torch.empty(B, P, max(k, 1), dtype=torch.int32, device="cuda")
dq = torch.empty(B, P, device="cuda"); iq = torch.empty(B, P, dtype=torch.int32, device="cuda")
kq = torch.empty(B, P, max(k, 1), dtype=torch.int32, device="cuda")
ws = torch.empty(_lib.load().mrb_knn_workspace_bytes(B, P, P), dtype=torch.uint8, device="cuda")
def run():
    _lib.call("mrb_knn_fwd", _lib.ptr(p), _lib.ptr(q), B, P, P, k, _lib.ptr(dp), _lib.ptr(ip), _lib.ptr(kp),
              _lib.ptr(dq), _lib.ptr(iq), _lib.ptr(kq), _lib.ptr(ws))
for _ in range(3): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
a.record()
for _ in range(n): run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / n
print("k=%d  %.3f ms/call (both directions)  %.3f Tpairs/s" % (k, ms, 2 * B * P * P / ms / 1e9))

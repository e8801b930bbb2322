"""Micro-benchmark of mrb_normals_fwd / mrb_normals_bwd_ld at the bench shape (B = 32, P = 10 k, k = 10; neighbour sets from
the k-NN of two surface clouds, like the step)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshrcnn_b200 import functional as F_, _lib, synthetic
from meshrcnn_b200.layers import Cubify
B, P, k = 32, 10000, 10
dev = "cuda"
v, vi, f, fi, _ = Cubify(0.2)(synthetic.blob_voxels(B, 24, 0).to(dev))
p, _ = F_.sample_points(v, f, vi, fi, P, seed=1)
q, _ = F_.sample_points(v + 0.3, f, vi, fi, P, seed=2)
_, _, kp, _, _, _ = F_.knn_search(p, q, k)
n = torch.empty(B, P, 3, device=dev); gn = torch.randn(B, P, 3, device=dev); g4 = torch.zeros(B, P, 4, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(run, reps=10):
    for _ in range(3): run()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / reps
fwd = timeit(lambda: _lib.call("mrb_normals_fwd", _lib.ptr(p), _lib.ptr(kp), B, P, k, _lib.ptr(n)))
bwd = timeit(lambda: _lib.call("mrb_normals_bwd_ld", _lib.ptr(p), _lib.ptr(kp), B, P, k, _lib.ptr(gn), _lib.ptr(g4), 4, None))
eig = torch.empty(12, B * P, dtype=torch.float64, device=dev)
fwd_e = timeit(lambda: _lib.call("mrb_normals_fwd_eig", _lib.ptr(p), _lib.ptr(kp), B, P, k, _lib.ptr(n), _lib.ptr(eig)))
bwd_e = timeit(lambda: _lib.call("mrb_normals_bwd_ld", _lib.ptr(p), _lib.ptr(kp), B, P, k, _lib.ptr(gn), _lib.ptr(g4), 4, _lib.ptr(eig)))
print("normals fwd %.1f us  bwd %.1f us | with the saved eigen-decomposition: fwd %.1f us  bwd %.1f us  (B=%d P=%d k=%d; |n| mean %.4f)"
      % (fwd * 1e3, bwd * 1e3, fwd_e * 1e3, bwd_e * 1e3, B, P, k, float(n.norm(dim=2).mean())))

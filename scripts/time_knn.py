"""Micro-benchmark of mrb_knn_fwd on bench-like clouds (B=32, P=Q=10k surface samples of two different blob sets,
the predicted one perturbed like a randomly initialised head) -- BASELINE config 5 size."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshrcnn_b200 import functional as F_, _lib, synthetic
from meshrcnn_b200.layers import Cubify
from meshrcnn_b200.mesh_sampling import normalize_mesh

B, P, k = 32, 10000, int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = "cuda"
v, vi, f, fi, _ = Cubify(0.2)(synthetic.blob_voxels(B, 24, 0).to(dev))
g = torch.Generator().manual_seed(0)
v = v + torch.rand(v.shape, generator=g).to(dev)            # tanh offsets of a random head are O(1) voxel units
p, _ = F_.sample_points(v, f, vi, fi, P, seed=1)
gv, gvi, gf, gfi, _ = Cubify(0.5)(synthetic.blob_voxels(B, 24, 1000).to(dev))
gt = torch.cat([normalize_mesh(x) for x in gv.split(gvi)])
q, _ = F_.sample_points(gt, gf, gvi, gfi, P, seed=2)
dp = torch.empty(B, P, device=dev); ip = torch.empty(B, P, dtype=torch.int32, device=dev)
kp = torch.empty(B, P, max(k, 1), dtype=torch.int32, device=dev)
dq = torch.empty(B, P, device=dev); iq = torch.empty(B, P, dtype=torch.int32, device=dev)
kq = torch.empty(B, P, max(k, 1), dtype=torch.int32, device=dev)
ws = torch.empty(_lib.load().mrb_knn_workspace_bytes(B, P, P), dtype=torch.uint8, device=dev)
algo = int(sys.argv[2]) if len(sys.argv) > 2 else 0        # 0 auto, 1 tiled scan, 2 cell grid (warp per query), 3 cell grid (thread per query)
def run():
    _lib.call("mrb_knn_fwd_algo", _lib.ptr(p), _lib.ptr(q), B, P, P, k, _lib.ptr(dp), _lib.ptr(ip), _lib.ptr(kp),
              _lib.ptr(dq), _lib.ptr(iq), _lib.ptr(kq), _lib.ptr(ws), algo)
for _ in range(3): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
a.record()
for _ in range(n): run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / n
print("algo=%d k=%d  %.3f ms/call (both directions)  %.3f Tpairs/s  mean NN d^2 = %.4f" % (algo, k, ms, 2 * B * P * P / ms / 1e9, float(dp.mean())))

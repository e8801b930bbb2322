"""cProfile view of the host side of one bench step (Python + ctypes + autograd dispatch), GPU work left asynchronous."""
import sys, os, time, cProfile, pstats, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from meshrcnn_b200.layers import Cubify
from meshrcnn_b200.mesh_sampling import normalize_mesh
from meshrcnn_b200.pipeline import MeshTargets, RefinementHead, weighted_loss
from meshrcnn_b200.sharding import FlatGradBucket
dev = torch.device("cuda", 0)
B = 32
vox_h, fmap_h, gt_vox_h = bench.make_inputs(B, 0)
sizes = [(224, 224)] * B
torch.manual_seed(1)
head = RefinementHead("pix3d", cubify_threshold=0.2).to(dev).train()
bucket = FlatGradBucket(head.parameters())
gv, gvi, gfaces, gfi, _ = Cubify(0.5)(gt_vox_h.to(dev))
gt = MeshTargets(torch.cat([normalize_mesh(v) for v in gv.split(gvi)]), gfaces, gvi, gfi)
vox_d = vox_h.to(dev); fmap_d = fmap_h.to(dev).requires_grad_()
def step():
    bucket.zero(); fmap_d.grad = None
    losses = head(vox_d, fmap_d, sizes, gt)
    weighted_loss(losses).backward()
for _ in range(5): step()
torch.cuda.synchronize()
N = 20
t0 = time.perf_counter()
for _ in range(N): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host time per step (launch side only): %.3f ms; incl. final sync: %.3f ms" % ((t1 - t0) / N * 1e3, (t2 - t0) / N * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(N): step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(35)

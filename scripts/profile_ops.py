"""torch.profiler view of one bench step: which aten ops (torch glue) still launch kernels on the hot path."""
import sys, os, torch
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
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))

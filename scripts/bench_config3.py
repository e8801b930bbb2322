"""BASELINE configs[2] per-GPU shard on one B200: residual ShapeNet head (ResVertixRefineShapenet x 3), 32 meshes from 48^3 blob
grids (~206k vertices), four ResNet50 maps of a 137 x 137 image (3840 channels), 10k-point losses, fwd + bwd.

    python scripts/bench_config3.py [model=shapenet_residual|shapenet] [steps]
"""
import json, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshrcnn_b200 import synthetic, _lib
from meshrcnn_b200.layers import Cubify
from meshrcnn_b200.mesh_sampling import normalize_mesh
from meshrcnn_b200.pipeline import MeshTargets, RefinementHead, weighted_loss
from meshrcnn_b200.sharding import FlatGradBucket

model = sys.argv[1] if len(sys.argv) > 1 else "shapenet_residual"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
B, V = 32, 48
vox = synthetic.blob_voxels(B, V, 0).to(dev)
fmaps = [m.to(dev).requires_grad_() for m in synthetic.feature_maps(B, synthetic.SHAPENET_MAPS, 0)]
sizes = [(137, 137)] * B
torch.manual_seed(1)
head = RefinementHead(model, cubify_threshold=0.2).to(dev).train()
bucket = FlatGradBucket(head.parameters())
gv, gvi, gf, gfi, _ = Cubify(0.5)(synthetic.blob_voxels(B, V, 1000).to(dev))
gt = MeshTargets(torch.cat([normalize_mesh(v) for v in gv.split(gvi)]), gf, gvi, gfi)

def step():
    bucket.zero()
    for m in fmaps:
        m.grad = None
    losses = head(vox, fmaps, sizes, gt)
    weighted_loss(losses).backward()
    return losses

for _ in range(3):
    losses = step()
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
for a, b in ev:
    a.record(); step(); b.record()
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in ev)
head.overlap_losses = False
with _lib.timed_calls() as tc:
    step()
v, vi, f, fi, adj = head.cubify(vox)
print(json.dumps({"config": "BASELINE configs[2], one GPU's shard: %s head, B=%d, %d^3 blobs" % (model, B, V), "SV": int(v.shape[0]),
                  "E": int(adj.shape[1]), "params": sum(p.numel() for p in head.parameters()), "ms_per_step_median": round(ms[len(ms) // 2], 3),
                  "ms_per_step_mean": round(sum(ms) / len(ms), 3), "meshes_per_s": round(B / (sum(ms) / len(ms)) * 1e3, 1),
                  "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2),
                  "losses": {k: float(x) for k, x in losses.items()},
                  "breakdown_ms": {k: round(x, 3) for k, x in sorted(tc.ms.items(), key=lambda kv: -kv[1])[:12]}}))

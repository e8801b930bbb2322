"""One step of a bench workload inside an NVTX range ("profiled"), after warm-up, for ncu:

    ncu --set full --nvtx --nvtx-include "profiled/" -o out python scripts/profile_step.py [pix3d|shapenet_residual|cubify4|bf16]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from meshrcnn_b200 import build, _lib, synthetic
build.build(); _lib.load()
what = sys.argv[1] if len(sys.argv) > 1 else "pix3d"
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
if what == "cubify4":
    from meshrcnn_b200.layers import Cubify
    vox = synthetic.dense_voxels(64, 48, 0).to(dev)
    cub = Cubify(0.5)
    run = lambda: cub(vox)
else:
    wl = bench.HeadWorkload("pix3d" if what == "bf16" else what, dev, 0, 1,
                            map_dtype=torch.bfloat16 if what == "bf16" else torch.float32)
    wl.head.overlap_losses = False          # single stream: per-kernel durations are not stretched by co-running kernels
    run = wl.step
for _ in range(4):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()                 # `ncu --profile-from-start off`: forward AND backward (the autograd thread is outside the NVTX range)
torch.cuda.nvtx.range_push("profiled")
run()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
torch.cuda.profiler.stop()

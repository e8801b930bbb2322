import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshrcnn_b200 import synthetic, _lib
from meshrcnn_b200.layers import Cubify
vox = synthetic.dense_voxels(64, 48, 0).cuda()
cub = Cubify(0.5)
for _ in range(3): cub(vox)
torch.cuda.synchronize()
with _lib.timed_calls() as tc:
    for _ in range(5): out = cub(vox)
print({k: round(v / 5, 3) for k, v in tc.ms.items()})

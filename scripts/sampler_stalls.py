"""Does the background `nvidia-smi -lms` clock sampler stall the timed steps?  60 bench steps per sampler setting."""
import sys, os, time, subprocess, shutil, torch
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
    weighted_loss(head(vox_d, fmap_d, sizes, gt)).backward()
for _ in range(5): step()
torch.cuda.synchronize()
import gc; gc.disable()
R = "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
NSTEP = int(sys.argv[1]) if len(sys.argv) > 1 else 60
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
FLUSH = False
for name, fields, ms in (("none", None, 0), ("flush only", "F", 0), ("flush + sampler @200ms", "F" + "clocks.sm,clocks.max.sm," + R, 200),
                         ("none again", None, 0), ("flush only again", "F", 0)):
    FLUSH = bool(fields) and fields.startswith("F")
    fields = fields[1:] if FLUSH else fields
    proc = None
    if fields:
        proc = subprocess.Popen([shutil.which("nvidia-smi"), "-i", "0", "--query-gpu=" + fields, "--format=csv,noheader,nounits", "-lms", str(ms)],
                                stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        time.sleep(1.0)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(NSTEP)]
    for a, b in ev:
        if FLUSH:
            flush.zero_()
        a.record(); step(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    med = t[len(t) // 2]
    print("%-26s median %.2f ms  max %.2f  steps > 1.3 x median: %d  mean %.3f  worst5 %s" % (name, med, t[-1], sum(x > 1.3 * med for x in t), sum(t) / len(t), ["%.1f" % x for x in t[-5:]]))
    if proc:
        proc.terminate(); proc.wait()

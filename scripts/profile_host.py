"""Host side of one headline step: issue time per step with the GPU idle at the start of every step (the e2e situation),
allocator activity per step, and a cProfile listing.      python scripts/profile_host.py [steps]"""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from meshrcnn_b200 import build, _lib
build.build(); _lib.load()
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
wl = bench.HeadWorkload("pix3d", dev, 0, 1)
for _ in range(20):
    wl.step()
torch.cuda.synchronize()

def host_times(fn, n):
    out = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); t1 = time.perf_counter()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        out.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
    out.sort()
    return out[len(out) // 2]

s0 = torch.cuda.memory_stats(dev)
print("host issue / total ms per step (median), staged buckets + hooks: %.3f / %.3f" % host_times(wl.step, N))
s1 = torch.cuda.memory_stats(dev)
print("cudaMalloc segments allocated during %d steps: %d; allocations per step: %.0f" % (
    N, s1["segment.all.allocated"] - s0["segment.all.allocated"], (s1["allocation.all.allocated"] - s0["allocation.all.allocated"]) / N))
import gc
gc.collect(); gc.disable()
print("same with gc disabled: %.3f / %.3f" % host_times(wl.step, N))
gc.enable()
wl.head.overlap_losses = False
print("single stream (no loss-stream overlap): %.3f / %.3f" % host_times(wl.step, N))
wl.head.overlap_losses = True
# count ATen vs own launches with the profiler-free method: torch.profiler is heavy; use _lib.launch_count for own kernels
l0 = _lib.launch_count; wl.step(); print("own kernel launches per step:", _lib.launch_count - l0)
pr = cProfile.Profile(); pr.enable()
for _ in range(N):
    wl.step()
pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(40)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    wl.step(); torch.cuda.synchronize()
ka = prof.key_averages()
rows = sorted([(k.key, k.count, getattr(k, "device_time_total", 0.0)) for k in ka if getattr(k, "device_time_total", 0) > 0 and k.count], key=lambda r: -r[2])
print("GPU kernels of one step (torch.profiler): name, launches, total us")
for r in rows[:60]:
    print("  %-90s %4d %9.1f" % (r[0][:90], r[1], r[2]))

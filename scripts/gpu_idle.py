"""Where the device waits for the launching thread: kernel timeline (torch.profiler / Kineto) of one step that starts with an
idle device (like an end-to-end step) and of one step issued right behind another (like the resident loop); prints the busy
time, the span and the largest gaps with the kernels around them."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from meshrcnn_b200 import build, _lib
build.build(); _lib.load()
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
wl = bench.HeadWorkload("pix3d", dev, 0, 1)
for _ in range(6):
    wl.step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity


def timeline(n_steps):
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(n_steps):
            wl.step()
        torch.cuda.synchronize()
    ev = [(e.time_range.start, e.time_range.end, e.name) for e in prof.events()
          if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    ev.sort()
    return ev


def report(ev, title):
    t0, t1 = ev[0][0], max(e[1] for e in ev)
    # union of busy intervals over all streams
    busy, cur_s, cur_e, gaps = 0.0, ev[0][0], ev[0][1], []
    last_name = ev[0][2]
    for s, e, name in ev[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, cur_e - t0, last_name, name))
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
        last_name = name
    busy += cur_e - cur_s
    print("%s: span %.0f us, device busy %.0f us, idle %.0f us in %d gaps" % (title, t1 - t0, busy, t1 - t0 - busy, len(gaps)))
    for g in sorted(gaps, reverse=True)[:8]:
        print("   gap %6.0f us at t = %6.0f us   after %-40s before %s" % (g[0], g[1], g[2][:40], g[3][:40]))
    small = sum(g[0] for g in gaps if g[0] < 10)
    print("   gaps < 10 us: %.0f us in total" % small)


one = timeline(1)
report(one, "one step from an idle device")
two = timeline(3)
# the middle step of three: from the first kernel launched after step 1's last ... approximate by thirds of the event list
n = len(two) // 3
report(two[n:2 * n], "the middle one of three back-to-back steps")

"""Root-cause hunt for the ~80 ms host stall of the e2e loop ("End-to-end path").

    python scripts/stall_hunt.py [old|new] [steps]

Runs the headline workload's e2e loop the round-1 way ("old": fresh timing events per loop, fresh `.to(non_blocking)`
device tensors, `.cpu()` read-back, 2 untimed steps) or the round-2 way ("new": bench.py's pre-recorded event pool, staging
buffers, pinned read-back, 12 untimed steps) with a watchdog thread that samples the main thread's Python stack every
2 ms.  For every step slower than 3 x the median it prints the stack samples that fall into that step (where the launching
thread was stuck) and the wall-clock gaps in which the watchdog itself could not run (GIL held by a C call).
"""
import collections, json, os, sys, threading, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

mode = sys.argv[1] if len(sys.argv) > 1 else "old"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
from meshrcnn_b200 import build, _lib
build.build(); _lib.load()
wl = bench.HeadWorkload("pix3d", dev, 0, 1)
sampler = bench.ClockSampler(0) if "--no-smi" not in sys.argv else None
if sampler:
    sampler.start(); time.sleep(0.5)
t0 = time.perf_counter()
while time.perf_counter() - t0 < 2.0:
    wl.step(exchange=False)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

main_id = threading.main_thread().ident
samples = collections.deque(maxlen=200000)      # (t, "file:line < file:line ...")
stop = False
def watchdog():
    while not stop:
        fr = sys._current_frames().get(main_id)
        if fr is not None:
            st = traceback.extract_stack(fr, limit=6)
            samples.append((time.perf_counter(), " < ".join("%s:%d" % (os.path.basename(f.filename), f.lineno) for f in reversed(st))))
        time.sleep(0.002)
th = threading.Thread(target=watchdog, daemon=True); th.start()

marks = []
if mode == "old":
    vox_pin, fmap_pin = wl.vox_h.pin_memory(), wl.fmaps_h[0].pin_memory()
    def one():
        v = vox_pin.to(dev, non_blocking=True)
        f = fmap_pin.to(dev, non_blocking=True).requires_grad_()
        losses = wl.step(v, [f])
        return torch.stack([losses["chamfer_loss"], losses["normal_loss"], losses["edge_loss"]]).detach().cpu()
    for _ in range(2):
        one()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    segs = []            # cudaMalloc'ed segments of the caching allocator after every step
    for a, b in ev:
        flush.zero_(); a.record(); ta = time.perf_counter(); one(); marks.append((ta, time.perf_counter())); b.record()
        segs.append(torch.cuda.memory_stats(dev)["segment.all.allocated"])
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
else:
    pool = bench.EventPool(steps + 8)
    sync = torch.cuda.synchronize
    orig = bench.timed_e2e
    # same code path as bench.py, with host marks around every step
    import types
    ms, phases, _ = bench.timed_e2e(wl, pool, steps, 12, flush, sync)
    tnow = time.perf_counter()
    marks = None
stop = True; th.join()
if sampler:
    print("clocks", sampler.stop())
med = sorted(ms)[len(ms) // 2]
print(json.dumps({"mode": mode, "steps": steps, "median_ms": round(med, 3), "max_ms": round(max(ms), 3), "mean_ms": round(sum(ms) / len(ms), 3),
                  "slow_steps": [(i, round(x, 2)) for i, x in enumerate(ms) if x > 1.3 * med]}))
if mode == "new":
    print("host phases of slow steps:", [(i, [round(p * 1e3, 1) for p in phases[i]]) for i, x in enumerate(ms) if x > 1.3 * med])
# watchdog gaps (GIL held / process descheduled) and the stacks inside slow steps
ts = [t for t, _ in samples]
gaps = [(ts[i] - ts[i - 1], samples[i - 1][1], samples[i][1]) for i in range(1, len(ts)) if ts[i] - ts[i - 1] > 0.02]
print("watchdog gaps > 20 ms:", len(gaps))
for g, before, after in gaps[:10]:
    print("  gap %.1f ms\n    before: %s\n    after:  %s" % (g * 1e3, before, after))
if marks:
    grew = [i for i in range(1, len(segs)) if segs[i] > segs[i - 1]]
    print("steps during which the caching allocator called cudaMalloc (new segments):", grew, "segments:", segs[0], "->", segs[-1])
    for i, x in enumerate(ms):
        if x > 3 * med:
            ta, tb = marks[i]
            inside = collections.Counter(s for t, s in samples if ta <= t <= tb)
            print("step %d: %.1f ms (host %.1f ms); stack samples:" % (i, x, (tb - ta) * 1e3))
            for s, c in inside.most_common(6):
                print("   %4d x %s" % (c, s))

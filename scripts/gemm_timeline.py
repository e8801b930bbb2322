"""Per-role timeline (clock64 stamps of CTA 0) of one k_gemm_tc launch; needs a -DMRB_TC_TIMELINE build of the library
(scripts/gemm_timeline.sh sets MRB_LIB_PATH).  Prints the merged event list in cycles relative to the first stamp."""
import sys, os, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshrcnn_b200 import functional as F_, _lib
wgrad = len(sys.argv) > 1 and sys.argv[1] == "wgrad"
if wgrad: sys.argv.pop(1)
M, K, N = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (50353, 128, 256)))
lib = ctypes.CDLL(os.environ["MRB_LIB_PATH"])
a = torch.randn(M, K, device="cuda"); w = torch.randn(K, N, device="cuda"); c = torch.empty(M, N, device="cuda")
img = F_.tc_pack(w, None, N, 1, 0, 0, K, N)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
run = lambda: F_.tc_gemm(_lib.ptr(a), K, M, K, img, N, _lib.ptr(c), N)
if wgrad:      # C[K x N] += X^T G, X = a [M, K], G = c-shaped [M, N]
    g = torch.randn(M, N, device="cuda"); gw = torch.zeros(K, N, device="cuda")
    run = lambda: _lib.call("mrb_gemm_tc_wgrad", _lib.ptr(a), K, _lib.ptr(g), N, M, K, N, _lib.ptr(gw), None, N, N)
for _ in range(3): run()
flush.zero_(); torch.cuda.synchronize()
stamps = (ctypes.c_longlong * (4 * 2048))(); counts = (ctypes.c_int * 4)()
lib.mrb_debug_tc_timeline(stamps, counts, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
lib.mrb_debug_tc_timeline(stamps, counts, 0)
names = {46: "A  all chunks stored, waiting for the MMAs", 50: "A  stage free", 51: "A  chunk stored", 40: "EP accumulator ready", 41: "EP tmem drained", 42: "EP tile stored",
         43: "EP   block in registers", 44: "EP   block in smem", 45: "EP   block stores issued",
         10: "MMA tile start (tmem free)", 20: "MMA chunk ready", 21: "MMA chunk issued", 60: "TMA stage free -> issue"}
ev = []
for r in range(4):
    for i in range(counts[r]):
        v = stamps[r * 2048 + i]
        ev.append((v >> 8, v & 255))
ev.sort()
t0 = ev[0][0]
print("M=%d K=%d N=%d  kernel %.1f us; CTA 0 events (cycles from first stamp):" % (M, K, N, e0.elapsed_time(e1) * 1e3))
only = os.environ.get("TL_ROLE")
for t, tag in ev:
    if only is None or names[tag].startswith(only):
        print("%8d  %s" % (t - t0, names[tag]))

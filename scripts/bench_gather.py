"""Micro-benchmark of the CSR gather kernels (old generic vs split-input v2) at the config-2 / config-3 sizes.
    python scripts/bench_gather.py [grid=48]"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meshrcnn_b200 import _lib, functional as F_, synthetic, build
from meshrcnn_b200.layers import Cubify
build.build(); _lib.load()
dev = torch.device("cuda", 0)
V = int(sys.argv[1]) if len(sys.argv) > 1 else 48
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
v, vi, f, fi, adj = Cubify(0.2)(synthetic.blob_voxels(32, V, 0).to(dev))
topo = F_.lookup(adj, v.shape[0])
SV, E, D = v.shape[0], adj.shape[1], 128
y = torch.randn(SV, 2 * D, device=dev); out = torch.empty(SV, D, device=dev); res = torch.randn(SV, D, device=dev)
mask = torch.empty(SV, D // 32, dtype=torch.int32, device=dev)
pos = v.clone(); w = torch.randn(2, 3, D, device=dev)
gout = torch.randn(SV, D, device=dev); gy = torch.empty(SV, 2 * D, device=dev); gpos = torch.empty(SV, 3, device=dev)
P = _lib.ptr

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(n):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b)
    return tot / n * 1e3
byts = 3 * 4 * SV * D + 4 * (E + SV + 1)
r = {"SV": SV, "E": E, "compulsory_bytes_fwd": byts}
r["old_fwd_relu_us"] = timeit(lambda: F_._gather(topo.rowptr, topo.col, SV, P(y), 2 * D, P(y) + 4 * D, 2 * D, D, True, P(out), D))
def new_fwd(pos_=None, mask_=None, res_=None):
    _lib.call("mrb_gc_gather_fwd", P(topo.rowptr), P(topo.col), SV, D, P(y), 2 * D, P(pos_), P(w) if pos_ is not None else None,
              P(w) + 4 * 3 * D if pos_ is not None else None, None, None, 1, P(mask_), P(res_), D if res_ is not None else 0, P(out), D)
r["new_fwd_plain_us"] = timeit(lambda: new_fwd())
r["new_fwd_mask_us"] = timeit(lambda: new_fwd(mask_=mask))
r["new_fwd_mask_res_us"] = timeit(lambda: new_fwd(mask_=mask, res_=res))
r["new_fwd_mask_pos_us"] = timeit(lambda: new_fwd(pos_=pos, mask_=mask))
act = torch.relu(torch.randn(SV, D, device=dev))
r["old_bwd_us"] = timeit(lambda: _lib.call("mrb_graphconv_bwd_gather", P(topo.rowptr_t), P(topo.col_t), SV, P(gout), D, P(act), D, D, P(gy)))
new_fwd(mask_=mask)
r["new_bwd_us"] = timeit(lambda: _lib.call("mrb_gc_gather_bwd", P(topo.rowptr_t), P(topo.col_t), SV, D, P(gout), D, P(mask), P(gy), None, None, None, None, None, 0))
r["new_bwd_gpos_us"] = timeit(lambda: _lib.call("mrb_gc_gather_bwd", P(topo.rowptr_t), P(topo.col_t), SV, D, P(gout), D, P(mask), P(gy), P(w), P(w) + 4 * 3 * D, P(gpos), None, None, 0))
for k in list(r):
    if k.endswith("_us"):
        r[k.replace("_us", "_GBps")] = round(byts / r[k] / 1e3, 1)
print(json.dumps(r, indent=1))

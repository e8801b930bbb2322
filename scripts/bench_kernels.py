"""Per-kernel roofline micro-benchmarks at the BASELINE.json config sizes (SURVEY.md section 8d figures).

    python scripts/bench_kernels.py > profiles/kernels_rXX.json

Each entry: algorithmic bytes (or pairs / flops) per launch, CUDA-event time (mean of N launches after warm-up, L2
flushed between launches for the HBM-bound kernels), achieved rate and fraction of the measured peak.
"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meshrcnn_b200 import _lib, functional as F_, synthetic
from meshrcnn_b200.layers import Cubify, VertexAlign

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
HBM = PEAKS["hbm_gbs"]
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=10, warm=3, l2_flush=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        if l2_flush:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n * 1e-3


out = {}

# ---- Cubify, config 4: 64 dense 48^3 grids ------------------------------------------------------------------
vox = synthetic.dense_voxels(64, 48, 0).to(dev)
cub = Cubify(0.5)
v, vi, f, fi, adj = cub(vox)
SV, SF, E = v.shape[0], f.shape[0], adj.shape[1]
byts = 4 * vox.numel() + 12 * SV + 24 * SF + 16 * E + 16 * 64
t = timeit(lambda: cub(vox), n=5)
with _lib.timed_calls() as tc:
    for _ in range(5):
        cub(vox)
out["cubify_config4"] = {"SV": SV, "SF": SF, "E": E, "bytes": byts, "ms": t * 1e3, "GBps": byts / t / 1e9, "frac_hbm": byts / t / 1e9 / HBM,
                         "ms_count_kernels": tc.ms["mrb_cubify_count"] / 5, "ms_emit_kernels": tc.ms["mrb_cubify_emit"] / 5,
                         "GBps_emit_only": (12 * SV + 24 * SF + 16 * E) / (tc.ms["mrb_cubify_emit"] / 5) / 1e6,
                         "note": "whole Cubify.forward incl. the count read-back sync and output allocation"}
del v, f, adj

# ---- Cubify, config 2 size (32 x 24^3 blobs) ----------------------------------------------------------------------
vox2 = synthetic.blob_voxels(32, 24, 0).to(dev)
cub2 = Cubify(0.2)
v, vi, f, fi, adj = cub2(vox2)
byts = 4 * vox2.numel() + 12 * v.shape[0] + 24 * f.shape[0] + 16 * adj.shape[1] + 16 * 32
t = timeit(lambda: cub2(vox2), n=10)
out["cubify_config2"] = {"bytes": byts, "ms": t * 1e3, "GBps": byts / t / 1e9, "frac_hbm": byts / t / 1e9 / HBM,
                         "note": "launch/sync latency bound at this size (6 small kernels + one D2H)"}

# ---- chamfer / kNN, config 5: B=32, P=Q=10k, surface-like clouds ---------------------------------------------------
B, P = 32, 10000
g = torch.Generator().manual_seed(0)
def cloud(seed):
    vv, vvi, ff, ffi, _ = Cubify(0.2)(synthetic.blob_voxels(B, 24, seed).to(dev))
    c, _ = F_.sample_points(vv * 0.05, ff, vvi, ffi, P, seed=seed + 1)
    return c
p, q = cloud(0), cloud(1000)
for algo in ("grid", "tiled"):
    for k in (0, 10):
        t = timeit(lambda: F_.knn_search(p, q, k, algo=algo), n=10, l2_flush=False)
        out["chamfer_knn_config5_k%d_%s" % (k, algo)] = {
            "pairs_per_call": 2 * B * P * P, "ms": t * 1e3, "Gpairs_per_s": 2 * B * P * P / t / 1e9,
            "frac_of_bruteforce_issue_bound": 2 * B * P * P / t / 1e12 / (148 * 128 * 1.965e9 / 1e12 / 4),
            "note": "one call = both directions (+ top-k index sets when k>0); 'grid' = exact cell-grid search (default), "
                    "'tiled' = shared-memory tiled scan with x pruning; bound = brute-force scan at 4 instr/pair"}

# ---- GraphConv pieces at config 3 per-GPU size (SV ~ 220k) -----------------------------------------------------------
v, vi, f, fi, adj = Cubify(0.2)(synthetic.blob_voxels(32, 48, 0).to(dev))
topo = F_.lookup(adj, v.shape[0])
SV, E = v.shape[0], adj.shape[1]
y = torch.randn(SV, 256, device=dev)
o = torch.empty(SV, 128, device=dev)
def gather():
    F_._gather(topo.rowptr, topo.col, SV, _lib.ptr(y), 256, _lib.ptr(y) + 512, 256, 128, True, _lib.ptr(o), 128)
t = timeit(gather)
byts = 4 * SV * 128 * 3 + 4 * (E + SV + 1)          # self + neighbour matrix read once + output
out["csr_gather_relu_config3"] = {"SV": SV, "E": E, "bytes_compulsory": byts, "ms": t * 1e3, "GBps": byts / t / 1e9,
                                  "frac_hbm": byts / t / 1e9 / HBM, "bytes_no_reuse": 4 * 128 * (E + 2 * SV)}
TF32_PEAK = PEAKS.get("bf16_tflops", 2250.0) / 2          # dense TF32 = half the measured bf16 rate
r4 = lambda v: (v + 3) // 4 * 4
for (K, N) in ((131, 256), (259, 256), (387, 256), (256, 131), (256, 387)):
    for pad in (False, True):
        lda, ldc = (r4(K), r4(N)) if pad else (K, N)
        if pad and lda == K and ldc == N:
            continue
        a = torch.randn(SV, lda, device=dev); w = torch.randn(K, N, device=dev); c = torch.empty(SV, ldc, device=dev)
        img = F_.tc_pack(w, None, N, 1, 0, 0, K, N)
        t = timeit(lambda: F_.tc_gemm(_lib.ptr(a), lda, SV, K, img, N, _lib.ptr(c), ldc))
        byts = 4 * SV * (K + N)
        out["gemm_tc_%dx%d_config3%s" % (K, N, "_rows16B" if pad else "")] = {
            "M": SV, "lda": lda, "ldc": ldc, "bytes": byts, "ms": t * 1e3, "GBps": byts / t / 1e9, "frac_hbm": byts / t / 1e9 / HBM,
            "TFLOPs_fp32_equiv": 2 * SV * K * N / t / 1e12, "TFLOPs_tf32_issued": 6 * SV * K * N / t / 1e12,
            "frac_tensor_tf32": 6 * SV * K * N / t / 1e12 / TF32_PEAK}
        del a, c
gy = torch.randn(SV, 256, device=dev)
for Kin in (131, 259, 387):
    x = torch.randn(SV, Kin, device=dev); gw = torch.zeros(2, Kin, 128, device=dev)
    def wgrad():
        gw.zero_()
        _lib.call("mrb_gemm_tc_wgrad", _lib.ptr(x), Kin, _lib.ptr(gy), 256, SV, Kin, 256, _lib.ptr(gw), _lib.ptr(gw) + 4 * Kin * 128, 128, 128)
    t = timeit(wgrad)
    byts = 4 * SV * (Kin + 256)
    out["gemm_tc_wgrad_%dx256_config3" % Kin] = {"bytes": byts, "ms": t * 1e3, "GBps": byts / t / 1e9, "frac_hbm": byts / t / 1e9 / HBM,
                                                 "TFLOPs_fp32_equiv": 2 * SV * Kin * 256 / t / 1e12,
                                                 "frac_tensor_tf32": 6 * SV * Kin * 256 / t / 1e12 / TF32_PEAK}
    del x
del gy

# ---- VertexAlign at config 3 (ShapeNet maps, 3840 channels) and config 2 (Pix3D) ---------------------------------------------
fm = [m.to(dev) for m in synthetic.feature_maps(32, synthetic.SHAPENET_MAPS, 0)]
pos = synthetic.in_frustum_positions(SV, 137, 0).to(dev)
al = VertexAlign().eval()
sizes = [(137, 137)] * 32
t = timeit(lambda: al(fm, pos, vi, sizes, [1] * 32), n=5)
byts = 4 * sum(m.numel() for m in fm) + 12 * SV + 4 * SV * 3840
out["vert_align_fwd_config3"] = {"bytes": byts, "ms": t * 1e3, "GBps": byts / t / 1e9, "frac_hbm": byts / t / 1e9 / HBM}

print(json.dumps(out, indent=1))

// Tiled brute-force nearest-neighbour / k-nearest-neighbour search between two batched point clouds, the
// chamfer sums, and the chamfer backward.
//
// Replaces batched_point2point_distance + batched_chamfer_distance + the two topk calls of compute_normals
// (reference meshRCNN/loss_functions.py:93-102,141,192-220), which materialise four dense B x P x Q fp32 tensors
// (1.6 GB per sample at P = Q = 10k).  Here nothing is materialised: every query point streams the other cloud
// through shared memory and keeps its running minimum / sorted top-k list in registers.  This kernel is bound
// by the FP32 CUDA-core issue rate (B*P*Q pairs, ~5.5 instructions each), not by HBM or the tensor pipe: inputs
// are 12 B/point and outputs <= 48 B/point.
//
// Distances are the direct sum of squared differences, which is *more* accurate than the reference's
// |p|^2 + |q|^2 - 2 p.q expansion (fp32 cancellation); parity is judged against the fp64 oracle.
#include "common.cuh"
#include "../../include/meshrcnn_b200.h"
#include <float.h>

namespace mrb {
namespace chamfer {

constexpr int TILE = 256;       // candidate points staged per shared-memory tile
constexpr int THREADS = 64;     // threads per CTA
constexpr int QPT = 1;          // query points per thread (one LDS.128 of a candidate feeds both)
constexpr int QCAP = 48;        // deferred-hit queue entries per query
constexpr int STEP = 32;         // candidates per group (32-bit hit mask) between two warp-wide queue checks
constexpr int SORT_MAX = 16384; // clouds up to this size are x-sorted in shared memory (pruned scan)

// ---------------------------------------------------------------------------------------------------------
// stage 0: per-cloud sort by the x coordinate (bitonic, one CTA per cloud) -> float4 (x, y, z, original index)
// ---------------------------------------------------------------------------------------------------------
constexpr int NBUCKET = 2048;

// One CTA per cloud: counting sort of the points into NBUCKET uniform x buckets (O(n): histogram with shared-memory
// atomics, block scan, scatter).  The order inside a bucket is arbitrary -- the scan kernel does not need a total order,
// only tiles (runs of TILE packed points) that are compact in x; their exact x ranges are written to `trange`.
// out: float4 (x, y, z, original index);  trange: [tile] (xmin, xmax).
__global__ void __launch_bounds__(1024) k_sort_x(const float* __restrict__ pts, int P, float4* __restrict__ out,
                                                 float2* __restrict__ trange, int ntiles) {
    extern __shared__ unsigned char smem_raw[];
    int* counts = reinterpret_cast<int*>(smem_raw);                 // [NBUCKET] histogram -> exclusive offsets -> cursors
    float* xs = reinterpret_cast<float*>(counts + NBUCKET);        // [P] x of the packed order
    __shared__ float s_lo[32], s_hi[32];
    __shared__ int s_scan[33];
    const float* src = pts + (size_t)blockIdx.x * P * 3;
    float lo = FLT_MAX, hi = -FLT_MAX;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const float x = src[3 * (size_t)i];
        lo = fminf(lo, x); hi = fmaxf(hi, x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane_id() == 0) { s_lo[warp_id()] = lo; s_hi[warp_id()] = hi; }
    for (int i = threadIdx.x; i < NBUCKET; i += blockDim.x) counts[i] = 0;
    __syncthreads();
    for (int w = 0; w < 32; ++w) { lo = fminf(lo, s_lo[w]); hi = fmaxf(hi, s_hi[w]); }
    const float scale = (hi > lo) ? (float)NBUCKET / (hi - lo) : 0.f;
    auto bucket_of = [&](float x) { return min(NBUCKET - 1, max(0, (int)((x - lo) * scale))); };
    for (int i = threadIdx.x; i < P; i += blockDim.x) atomicAdd(&counts[bucket_of(src[3 * (size_t)i])], 1);
    __syncthreads();
    // exclusive scan of the NBUCKET counts: 2 per thread
    {
        const int a = counts[2 * threadIdx.x], b2 = counts[2 * threadIdx.x + 1];
        int total;
        const int ex = block_exclusive_scan(a + b2, s_scan, &total);
        __syncthreads();
        counts[2 * threadIdx.x] = ex;
        counts[2 * threadIdx.x + 1] = ex + a;
    }
    __syncthreads();
    float4* dst = out + (size_t)blockIdx.x * P;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const float x = src[3 * (size_t)i];
        const int pos = atomicAdd(&counts[bucket_of(x)], 1);
        dst[pos] = make_float4(x, src[3 * (size_t)i + 1], src[3 * (size_t)i + 2], __int_as_float(i));
        xs[pos] = x;
    }
    __syncthreads();
    float2* tr = trange + (size_t)blockIdx.x * ntiles;
    for (int t = warp_id(); t < ntiles; t += blockDim.x >> 5) {
        float tlo = FLT_MAX, thi = -FLT_MAX;
        for (int i = t * TILE + lane_id(); i < min((t + 1) * TILE, P); i += 32) { tlo = fminf(tlo, xs[i]); thi = fmaxf(thi, xs[i]); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            tlo = fminf(tlo, __shfl_xor_sync(0xffffffffu, tlo, o));
            thi = fmaxf(thi, __shfl_xor_sync(0xffffffffu, thi, o));
        }
        if (lane_id() == 0) tr[t] = make_float2(tlo, thi);
    }
}

// clouds too large for the shared-memory sort: pack in the original order (no pruning)
__global__ void k_pack(const float* __restrict__ pts, long long n_total, int P, float4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    out[i] = make_float4(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], __int_as_float((int)(i % P)));
}

// ---------------------------------------------------------------------------------------------------------
// stage 1: pruned, filtered brute-force scan
// ---------------------------------------------------------------------------------------------------------
template <int K>
struct TopK {
    float d[K];
    int i[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < K; ++s) { d[s] = FLT_MAX; i[s] = 0x7fffffff; }
    }
    // (distance, index) lexicographic order: sorted ascending, ties -> lower original index first
    static __device__ __forceinline__ bool after(float da, int ia, float db, int ib) {
        return da > db || (da == db && ia > ib);
    }
    __device__ __forceinline__ bool admits(float nd, int ni) const { return after(d[K - 1], i[K - 1], nd, ni); }
    __device__ __forceinline__ void insert(float nd, int ni) {
#pragma unroll
        for (int s = K - 1; s > 0; --s) {
            const bool shift = after(d[s - 1], i[s - 1], nd, ni);
            const bool here = after(d[s], i[s], nd, ni);
            const float td = shift ? d[s - 1] : (here ? nd : d[s]);
            const int ti = shift ? i[s - 1] : (here ? ni : i[s]);
            d[s] = td; i[s] = ti;
        }
        if (after(d[0], i[0], nd, ni)) { d[0] = nd; i[0] = ni; }
    }
};

// One query point: exact top-K list plus the filter threshold derived from it.
template <int K>
struct Query {
    float x, y, z, pp;     // coordinates and |p|^2
    float thr;             // filter threshold on s = |q|^2 - 2 p.q  (== d - |p|^2 up to rounding)
    int cnt;               // queued hits of the current tile
    TopK<K> top;
    __device__ __forceinline__ void set_threshold(float margin) { thr = top.d[K - 1] - pp + margin; }
    // exact re-evaluation of the queued candidates with the direct (p - q)^2 form
    __device__ __forceinline__ void drain(const float4* __restrict__ tile, const int* __restrict__ tile_idx,
                                          const unsigned short* __restrict__ queue, float margin) {
        for (int j = 0; j < cnt; ++j) {
            const int t = queue[j * THREADS];
            const float4 c = tile[t];
            // c = (-2qx, -2qy, -2qz, |q|^2): the coordinates are recovered exactly (power-of-two scaling)
            const float dx = x + 0.5f * c.x, dy = y + 0.5f * c.y, dz = z + 0.5f * c.z;
            const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const int ci = tile_idx[t];
            if (top.admits(d, ci)) top.insert(d, ci);
        }
        cnt = 0;
        set_threshold(margin);
    }
};

// grid: (ceil(P / (THREADS*QPT)), B).  `a` / `b` are the packed (x-sorted when prune != 0) clouds of stage 0.
// For every point of a: squared distance + original index of the nearest point of b, and the K nearest indices sorted
// by (distance, index); results are written at the query's original index.
//
// * Hot loop per candidate and query: 3 FFMA (expanded form |q|^2 - 2 p.q against a per-query threshold) + 1 compare --
//   a conservative *filter* whose rounding error is covered by `margin`; hits are queued (2 bytes) and re-evaluated
//   exactly when a queue fills (warp-wide) or the tile ends, so results equal an exact scan while the divergent
//   sorted-list insertion stays out of the inner loop.
// * Pruning: both clouds are bucket-sorted by x, a CTA's queries span a narrow x range, candidate tiles are visited outwards
//   from that range, and a direction is abandoned once the tile's x gap alone exceeds every query's current k-th best
//   distance.  The scan stays exact (the bound is a true lower bound of the distance).
template <int K>
__global__ void __launch_bounds__(THREADS) k_nn(const float4* __restrict__ a, const float4* __restrict__ b,
                                                const float2* __restrict__ trange_b, int P, int Q, int prune,
                                                float* __restrict__ min_d, int32_t* __restrict__ min_i,
                                                int32_t* __restrict__ knn, int k_out) {
    __shared__ float4 tile[TILE + STEP];
    __shared__ int tile_idx[TILE];
    __shared__ unsigned short queue[QPT][QCAP * THREADS];
    __shared__ int qq_max_bits;
    __shared__ float red[THREADS / 32][3];
    __shared__ float blk[3];    // x range of the CTA's queries, max k-th best distance
    const int batch = blockIdx.y;
    const int tid = threadIdx.x;
    const float4* bq = b + (size_t)batch * Q;

    Query<K> qr[QPT];
    int pidx[QPT], orig[QPT];
    float xlo = FLT_MAX, xhi = -FLT_MAX;
#pragma unroll
    for (int u = 0; u < QPT; ++u) {
        pidx[u] = (blockIdx.x * QPT + u) * THREADS + tid;
        const float4 ap = a[(size_t)batch * P + min(pidx[u], P - 1)];   // out-of-range slots duplicate the last query
        qr[u].x = ap.x; qr[u].y = ap.y; qr[u].z = ap.z;
        orig[u] = __float_as_int(ap.w);
        qr[u].pp = fmaf(ap.z, ap.z, fmaf(ap.y, ap.y, ap.x * ap.x));
        qr[u].cnt = 0;
        qr[u].top.init();
        xlo = fminf(xlo, ap.x); xhi = fmaxf(xhi, ap.x);
    }
    // CTA-wide x range of the queries
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xlo = fminf(xlo, __shfl_xor_sync(0xffffffffu, xlo, o));
        xhi = fmaxf(xhi, __shfl_xor_sync(0xffffffffu, xhi, o));
    }
    if (lane_id() == 0) { red[warp_id()][0] = xlo; red[warp_id()][1] = xhi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < THREADS / 32; ++w) { xlo = fminf(xlo, red[w][0]); xhi = fmaxf(xhi, red[w][1]); }
        blk[0] = xlo; blk[1] = xhi; blk[2] = FLT_MAX;
    }
    __syncthreads();
    xlo = blk[0]; xhi = blk[1];

    const int ntiles = (Q + TILE - 1) / TILE;
    const float2* tr = trange_b + (size_t)batch * ntiles;
    // first tile: the first one whose x range reaches the query range
    int start = 0;
    if (prune) {
        while (start + 1 < ntiles && tr[start].y < xlo) ++start;
    }
    int left = start - 1, right = start;       // next unvisited tile on each side
    bool left_open = prune != 0, right_open = true;
    if (!prune) left = -1;

    while (true) {
        // ---- pick the next tile: the side whose x gap to the query range is smaller; prune by the gap --------------
        int tsel = -1;
        {
            const float thr_max = blk[2];
            float gl = FLT_MAX, gr = FLT_MAX;
            if (left_open && left >= 0) {
                const float xl = tr[left].y;                                // largest x of the left tile
                gl = fmaxf(xlo - xl, 0.f);
                if (thr_max < FLT_MAX && gl * gl > thr_max) { left_open = false; gl = FLT_MAX; }
            } else left_open = false;
            if (right_open && right < ntiles) {
                const float xr = prune ? tr[right].x : 0.f;                 // smallest x of the right tile
                gr = prune ? fmaxf(xr - xhi, 0.f) : 0.f;
                if (thr_max < FLT_MAX && gr * gr > thr_max) { right_open = false; gr = FLT_MAX; }
            } else right_open = false;
            if (!left_open && !right_open) break;
            if (right_open && (!left_open || gr <= gl)) tsel = right++; else tsel = left--;
        }
        const int q0 = tsel * TILE;
        const int cnt = min(TILE, Q - q0);
        __syncthreads();
        if (tid == 0) qq_max_bits = 0;
        __syncthreads();
        float qmax = 0.f;
        for (int t = tid; t < cnt; t += THREADS) {
            const float4 s = bq[q0 + t];
            const float qq = fmaf(s.z, s.z, fmaf(s.y, s.y, s.x * s.x));
            tile[t] = make_float4(-2.f * s.x, -2.f * s.y, -2.f * s.z, qq);
            tile_idx[t] = __float_as_int(s.w);
            qmax = fmaxf(qmax, qq);
        }
        if (tid < STEP) tile[cnt + tid] = make_float4(0.f, 0.f, 0.f, FLT_MAX);   // never-matching sentinels
        atomicMax(&qq_max_bits, __float_as_int(qmax));     // non-negative floats order like their bit patterns
        __syncthreads();
        const float qq_max = __int_as_float(qq_max_bits);
        float margin[QPT];
#pragma unroll
        for (int u = 0; u < QPT; ++u) {
            margin[u] = 1e-6f * (qr[u].pp + qq_max) + 1e-37f;
            qr[u].set_threshold(margin[u]);
        }
        // STEP candidates per iteration, then one warp-uniform vote: queues are drained by the whole warp together
        // (a lane draining alone would serialise the ~70-instruction insertion across the 32 lanes).
        for (int t0 = 0; t0 < cnt; t0 += STEP) {
            unsigned hit[QPT];
#pragma unroll
            for (int u = 0; u < QPT; ++u) hit[u] = 0u;
#pragma unroll
            for (int tt = 0; tt < STEP; ++tt) {
                const float4 c = tile[t0 + tt];
#pragma unroll
                for (int u = 0; u < QPT; ++u) {
                    const float s = fmaf(qr[u].x, c.x, fmaf(qr[u].y, c.y, fmaf(qr[u].z, c.z, c.w)));
                    if (s < qr[u].thr) hit[u] |= 1u << tt;             // one predicated LOP per pair, no address math
                }
            }
            int fullest = 0;
#pragma unroll
            for (int u = 0; u < QPT; ++u) {
                unsigned m = hit[u];
                while (m) {                                              // usually zero or one bit
                    queue[u][qr[u].cnt * THREADS + tid] = (unsigned short)(t0 + __ffs(m) - 1);
                    ++qr[u].cnt;
                    m &= m - 1;
                }
                fullest = max(fullest, qr[u].cnt);
            }
            if (__any_sync(0xffffffffu, fullest > QCAP - STEP)) {
#pragma unroll
                for (int u = 0; u < QPT; ++u) qr[u].drain(tile, tile_idx, &queue[u][tid], margin[u]);
            }
        }
        float tmax = 0.f;
#pragma unroll
        for (int u = 0; u < QPT; ++u) {
            qr[u].drain(tile, tile_idx, &queue[u][tid], margin[u]);
            tmax = fmaxf(tmax, qr[u].top.d[K - 1]);
        }
        // CTA-wide max of the k-th best distances (the pruning bound for the next tile)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        if (lane_id() == 0) red[warp_id()][2] = tmax;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < THREADS / 32; ++w) tmax = fmaxf(tmax, red[w][2]);
            blk[2] = tmax;
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < QPT; ++u) {
        if (pidx[u] >= P) continue;
        const size_t o = (size_t)batch * P + orig[u];
        min_d[o] = qr[u].top.d[0];
        min_i[o] = qr[u].top.i[0];
#pragma unroll
        for (int s = 0; s < K; ++s)
            if (s < k_out) knn[o * k_out + s] = qr[u].top.i[s];
    }
}

// ---------------------------------------------------------------------------------------------------------
// grid path: exact k-NN through a uniform cell grid over the candidate cloud (default for clouds <= 65535 points)
// ---------------------------------------------------------------------------------------------------------
// The tiled scan above visits ~19 % of all pairs on the refinement workload (the x gap is a weak bound for points on a
// surface).  Here each cloud is counting-sorted into G^3 cells (x fastest), so the points of cells [x0, x1] of one
// (z, y) row are ONE contiguous run of the sorted array.  A query scans the rows of the cell box that covers
// [q - r, q + r]; if the k-th best distance found is <= r^2 the result is final (every closer point lies inside the
// box), otherwise r becomes that k-th distance (or doubles while fewer than k points were seen) and only the new
// shell of cells is scanned.  ~60 candidates per query instead of ~1900, and the same (distance, index) order as the
// brute-force scan, so both paths return identical results.
//
// Exactness under rounding: cell_of() is a monotone non-decreasing fp32 function (rn subtract, rn multiply by a
// positive constant, floor, clamp) used for candidates and box corners alike, so a candidate whose coordinates lie in
// [q - rr, q + rr] can never fall outside the cell box; rr exceeds r by 1e-5 relative, far above the ~4e-7 relative
// rounding of the fp32 distance it is compared with.
constexpr int GMAX = 32;                 // cells per axis (G^3 int counters must fit in shared memory)
constexpr int GRID_MAX_POINTS = 65535;   // cell offsets are uint16
constexpr int GRID_THREADS = 128;
// (re-measured at the end of r2, scripts/variants.sh chamfer.cu "python scripts/time_knn.py 10 2": C0 3 / 4 / 5: 0.656 / 0.645 /
//  0.672 ms; points per cell column (DENS) 8 / 12 / 16 / 24: 0.698 / 0.645 / 0.655 / 0.663 ms -- the defaults stay)
#ifndef MRB_GRID_C0
#define MRB_GRID_C0 4.0f
#endif
#ifndef MRB_GRID_DENS
#define MRB_GRID_DENS 12.0
#endif
constexpr float GRID_C0 = MRB_GRID_C0;   // first box sized for ~GRID_C0 * k points (measured best of 2..8 on B200)
constexpr int CS_STRIDE = GMAX * GMAX * GMAX + 2;

struct __align__(16) GridHdr {
    float lo[3];
    float r0;       // first search radius
    float inv[3];   // cells per unit length (0 for a degenerate axis)
    int G;
};

__device__ __forceinline__ int cell_of(float x, float lo, float inv, int G) {
    return min(G - 1, max(0, __float2int_rd(__fmul_rn(__fsub_rn(x, lo), inv))));
}

static int grid_cells(int n) {
    int g = (int)ceil(sqrt((double)n / MRB_GRID_DENS));
    return g < 1 ? 1 : (g > GMAX ? GMAX : g);
}

// One CTA per cloud (clouds of a first, then clouds of b): bounding box, cell histogram in shared memory, exclusive scan
// (-> uint16 cell offsets), scatter into float4 (x, y, z, original index).  The order inside a cell is arbitrary; the
// search result does not depend on it.  r0_scale = sqrt(c0 * k): the first box is sized to hold ~c0 * k points if the
// cloud is a surface (occupied cells ~ box side^2).
__global__ void __launch_bounds__(1024) k_grid_build(const float* __restrict__ pa, const float* __restrict__ pb, int B, int P,
                                                     int Q, int Ga, int Gb, float4* __restrict__ oa, float4* __restrict__ ob,
                                                     uint16_t* __restrict__ csa, uint16_t* __restrict__ csb,
                                                     GridHdr* __restrict__ ha, GridHdr* __restrict__ hb, float r0_scale) {
    extern __shared__ unsigned char smem_raw[];
    int* counts = reinterpret_cast<int*>(smem_raw);
    __shared__ float s_red[32][6];
    __shared__ int s_scan[33];
    const bool second = (int)blockIdx.x >= B;
    const int cloud = second ? blockIdx.x - B : blockIdx.x;
    const int n = second ? Q : P, G = second ? Gb : Ga;
    const float* src = (second ? pb : pa) + (size_t)cloud * n * 3;
    float4* dst = (second ? ob : oa) + (size_t)cloud * n;
    uint16_t* cs = (second ? csb : csa) + (size_t)cloud * CS_STRIDE;
    GridHdr* hdr = (second ? hb : ha) + cloud;
    const int ncell = G * G * G;

    // Clouds of <= GB_PPT * 1024 points (the 10 000-point loss clouds) are read from global memory ONCE: every thread keeps
    // its points in registers for the three passes (bounding box, histogram, scatter), with all loads in flight together.
    constexpr int GB_PPT = 12;
    const bool in_regs = n <= GB_PPT * (int)blockDim.x;
    float rx[GB_PPT], ry[GB_PPT], rz[GB_PPT];
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (in_regs) {
#pragma unroll
        for (int u = 0; u < GB_PPT; ++u) {
            const int i = min(threadIdx.x + u * (int)blockDim.x, n - 1);      // clamped: duplicates do not move the box
            rx[u] = src[3 * (size_t)i]; ry[u] = src[3 * (size_t)i + 1]; rz[u] = src[3 * (size_t)i + 2];
        }
#pragma unroll
        for (int u = 0; u < GB_PPT; ++u) {
            lo[0] = fminf(lo[0], rx[u]); hi[0] = fmaxf(hi[0], rx[u]);
            lo[1] = fminf(lo[1], ry[u]); hi[1] = fmaxf(hi[1], ry[u]);
            lo[2] = fminf(lo[2], rz[u]); hi[2] = fmaxf(hi[2], rz[u]);
        }
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float v = src[3 * (size_t)i + d];
                lo[d] = fminf(lo[d], v); hi[d] = fmaxf(hi[d], v);
            }
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
        if (lane_id() == 0) { s_red[warp_id()][d] = lo[d]; s_red[warp_id()][3 + d] = hi[d]; }
    }
    for (int i = threadIdx.x; i < ncell; i += blockDim.x) counts[i] = 0;
    __syncthreads();
    float inv[3], hsum = 0.f;
    int hdims = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        for (int w = 0; w < 32; ++w) { lo[d] = fminf(lo[d], s_red[w][d]); hi[d] = fmaxf(hi[d], s_red[w][3 + d]); }
        const float ext = hi[d] - lo[d];
        inv[d] = (ext > 0.f && ext < FLT_MAX) ? (float)G / ext : 0.f;
        if (inv[d] > 0.f) { hsum += ext / (float)G; ++hdims; }
    }
    auto cell = [&](int i) {
        const int cx = cell_of(src[3 * (size_t)i], lo[0], inv[0], G);
        const int cy = cell_of(src[3 * (size_t)i + 1], lo[1], inv[1], G);
        const int cz = cell_of(src[3 * (size_t)i + 2], lo[2], inv[2], G);
        return (cz * G + cy) * G + cx;
    };
    auto cell_xyz = [&](float x, float y, float z) {
        return (cell_of(z, lo[2], inv[2], G) * G + cell_of(y, lo[1], inv[1], G)) * G + cell_of(x, lo[0], inv[0], G);
    };
    if (in_regs) {
#pragma unroll
        for (int u = 0; u < GB_PPT; ++u)
            if (threadIdx.x + u * (int)blockDim.x < n) atomicAdd(&counts[cell_xyz(rx[u], ry[u], rz[u])], 1);
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&counts[cell(i)], 1);
    }
    __syncthreads();
    const int per = (ncell + (int)blockDim.x - 1) / (int)blockDim.x;
    const int c0 = min(ncell, (int)threadIdx.x * per), c1 = min(ncell, c0 + per);
    int local = 0, occupied = 0;
    for (int c = c0; c < c1; ++c) { local += counts[c]; occupied += counts[c] > 0; }
    int total;
    int run = block_exclusive_scan(local, s_scan, &total);
    for (int c = c0; c < c1; ++c) {
        const int v = counts[c];
        counts[c] = run;
        cs[c] = (uint16_t)run;
        run += v;
    }
    __shared__ int s_occ[33];
    occupied = block_sum<int>(occupied, s_occ);
    if (threadIdx.x == 0) {
        cs[ncell] = (uint16_t)n;
        GridHdr h;
        const float hmean = hdims ? hsum / (float)hdims : 0.f;
        h.r0 = fmaxf(0.5f * hmean * r0_scale * sqrtf((float)max(occupied, 1) / (float)max(n, 1)), 1e-20f);
#pragma unroll
        for (int d = 0; d < 3; ++d) { h.lo[d] = lo[d]; h.inv[d] = inv[d]; }
        h.G = G;
        *hdr = h;
    }
    __syncthreads();
    if (in_regs) {
#pragma unroll
        for (int u = 0; u < GB_PPT; ++u) {
            const int i = threadIdx.x + u * (int)blockDim.x;
            if (i < n) {
                const int pos = atomicAdd(&counts[cell_xyz(rx[u], ry[u], rz[u])], 1);
                dst[pos] = make_float4(rx[u], ry[u], rz[u], __int_as_float(i));
            }
        }
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int pos = atomicAdd(&counts[cell(i)], 1);
            dst[pos] = make_float4(src[3 * (size_t)i], src[3 * (size_t)i + 1], src[3 * (size_t)i + 2], __int_as_float(i));
        }
    }
}

constexpr int GRID_QPW = 4;   // queries per warp (amortises the grid header)
constexpr int GRID_CH = 3;    // candidate rounds (32 each) gathered per selection

__device__ __forceinline__ bool key_less(float da, int ia, float db, int ib) { return da < db || (da == db && ia < ib); }
// orders two (distance, index) keys so that a <= b
__device__ __forceinline__ void key_sort2(float& da, int& ia, float& db, int& ib) {
    const bool sw = key_less(db, ib, da, ia);
    const float td = sw ? db : da;
    const int ti = sw ? ib : ia;
    db = sw ? da : db; ib = sw ? ia : ib;
    da = td; ia = ti;
}

// One WARP per query (GRID_QPW consecutive queries per warp), so nothing diverges and candidate loads are coalesced runs:
//   * lane = row of the cell box: loads the row's [start, end) from the cell offsets; a warp scan turns the row lengths
//     into one flat candidate numbering;
//   * lane = candidate: binary search (shuffles) for its row, one float4 load, exact (p - q)^2 distance, dropped unless it
//     beats the current k-th key;
//   * selection: each lane sorts its <= GRID_CH new keys + its slot of the running list, then the K smallest keys of the
//     warp are extracted with two redux.sync.min per key (distance bits, then index among the ties); list slot s lives in
//     lane s.
// A pass is final when K keys were found and the k-th distance is <= r^2; otherwise r grows to that distance (or doubles)
// and the search restarts with the old k-th distance as a filter.  grid: (ceil(P / (4 * GRID_QPW)), B).
#ifndef MRB_GRID_SKIP_UNDERFULL
#define MRB_GRID_SKIP_UNDERFULL 1
#endif
#ifndef MRB_GRID_MINB
#define MRB_GRID_MINB 10    // 48 registers (measured: 1 -> 78 regs 0.77 ms, 10 -> 0.67, 12 -> 0.65, 16 -> 0.66 ms per call at config 5)
#endif
template <int K>
__global__ void __launch_bounds__(GRID_THREADS, MRB_GRID_MINB) k_nn_grid(const float4* __restrict__ a, const float4* __restrict__ b,
                                                          const uint16_t* __restrict__ cs_b, const GridHdr* __restrict__ hdr_b,
                                                          int P, int Q, float* __restrict__ min_d, int32_t* __restrict__ min_i,
                                                          int32_t* __restrict__ knn, int k_out) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int NONE = 0x7fffffff;
    const int batch = blockIdx.y;
    const int lane = lane_id();
    const int wq0 = (blockIdx.x * (GRID_THREADS / 32) + warp_id()) * GRID_QPW;
    if (wq0 >= P) return;
    const GridHdr h = hdr_b[batch];
    const int G = h.G;
    const uint16_t* __restrict__ cs = cs_b + (size_t)batch * CS_STRIDE;
    const float4* __restrict__ bq = b + (size_t)batch * Q;
    for (int u = 0; u < GRID_QPW; ++u) {
        const int qi = wq0 + u;
        if (qi >= P) break;
        const float4 ap = a[(size_t)batch * P + qi];
        float ld = FLT_MAX;      // running list: slot `lane` (ascending over lanes 0..K-1)
        int li = NONE;
        float thr_d = FLT_MAX;   // only keys below (thr_d, thr_i) can still enter the list
        int thr_i = NONE;
        float r = h.r0;
        for (int pass = 0; pass < 512; ++pass) {
            const float rr = fmaf(r, 1e-5f, r) + 1e-30f;
            const int x0 = cell_of(ap.x - rr, h.lo[0], h.inv[0], G), x1 = cell_of(ap.x + rr, h.lo[0], h.inv[0], G);
            const int y0 = cell_of(ap.y - rr, h.lo[1], h.inv[1], G), y1 = cell_of(ap.y + rr, h.lo[1], h.inv[1], G);
            const int z0 = cell_of(ap.z - rr, h.lo[2], h.inv[2], G), z1 = cell_of(ap.z + rr, h.lo[2], h.inv[2], G);
            const int ny = y1 - y0 + 1, nrow = ny * (z1 - z0 + 1);
            const float inv_ny = 1.0f / (float)ny;
            const bool all = (h.inv[0] == 0.f || (x0 == 0 && x1 == G - 1)) && (h.inv[1] == 0.f || (y0 == 0 && y1 == G - 1)) &&
                             (h.inv[2] == 0.f || (z0 == 0 && z1 == G - 1));
            ld = FLT_MAX; li = NONE;
            for (int rbase = 0; rbase < nrow; rbase += 32) {
                int s = 0, cnt = 0;
                if (rbase + lane < nrow) {
                    const int rz = (int)(((float)(rbase + lane) + 0.5f) * inv_ny), ry = (rbase + lane) - rz * ny;   // exact: < 2^10 rows
                    const int row = ((z0 + rz) * G + (y0 + ry)) * G;
                    s = cs[row + x0];
                    cnt = (int)cs[row + x1 + 1] - s;
                }
                const int incl = warp_inclusive_scan(cnt);
                const int m = __shfl_sync(FULL, incl, 31);
#if MRB_GRID_SKIP_UNDERFULL
                // The cell offsets already tell how many candidates the box holds: a box (all its rows fit this batch) with
                // fewer than K of them cannot fill the list, so the pass could not be final -- grow r without reading a point.
                if (nrow <= 32 && m < K && !all) break;
#endif
                const int shift = s - (incl - cnt);                // candidate t of this row sits at bq[t + shift]
                for (int base = 0; base < m; base += 32 * GRID_CH) {
                    float vd[GRID_CH + 1];
                    int vi[GRID_CH + 1];
                    bool fresh = false;
#pragma unroll
                    for (int c = 0; c < GRID_CH; ++c) {
                        vd[c] = FLT_MAX; vi[c] = NONE;
                        if (base + c * 32 >= m) continue;           // warp-uniform: round not needed
                        const int t = base + c * 32 + lane;
                        int pos = 0;                                // number of rows that end at or before t
#pragma unroll
                        for (int step = 16; step > 0; step >>= 1) {
                            const int v = __shfl_sync(FULL, incl, pos + step - 1);
                            if (v <= t) pos += step;
                        }
                        const int sh = __shfl_sync(FULL, shift, pos);
                        if (t < m) {
                            const float4 cp = bq[t + sh];
                            const float dx = ap.x - cp.x, dy = ap.y - cp.y, dz = ap.z - cp.z;
                            const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                            const int ci = __float_as_int(cp.w);
                            if (key_less(d, ci, thr_d, thr_i)) { vd[c] = d; vi[c] = ci; fresh = true; }
                        }
                    }
                    if (!__any_sync(FULL, fresh)) continue;
                    vd[GRID_CH] = ld; vi[GRID_CH] = li;
                    static_assert(GRID_CH == 3, "sorting network below is for 4 keys");
                    key_sort2(vd[0], vi[0], vd[1], vi[1]);
                    key_sort2(vd[2], vi[2], vd[3], vi[3]);
                    key_sort2(vd[0], vi[0], vd[2], vi[2]);
                    key_sort2(vd[1], vi[1], vd[3], vi[3]);
                    key_sort2(vd[1], vi[1], vd[2], vi[2]);
                    ld = FLT_MAX; li = NONE;
#pragma unroll
                    for (int sidx = 0; sidx < K; ++sidx) {
                        const unsigned fb = __float_as_uint(vd[0]);
                        const unsigned mn = __reduce_min_sync(FULL, fb);
                        const unsigned mi = __reduce_min_sync(FULL, fb == mn ? (unsigned)vi[0] : 0xffffffffu);
                        if (lane == sidx) { ld = __uint_as_float(mn); li = (int)mi; }
                        if (fb == mn && (unsigned)vi[0] == mi) {
#pragma unroll
                            for (int c = 0; c < GRID_CH; ++c) { vd[c] = vd[c + 1]; vi[c] = vi[c + 1]; }
                            vd[GRID_CH] = FLT_MAX; vi[GRID_CH] = NONE;
                        }
                    }
                    const float kd = __shfl_sync(FULL, ld, K - 1);
                    const int ki = __shfl_sync(FULL, li, K - 1);
                    if (ki != NONE) { thr_d = kd; thr_i = ki; }
                }
            }
            const float kd = __shfl_sync(FULL, ld, K - 1);
            const bool full = __shfl_sync(FULL, li, K - 1) != NONE;
            if (full && kd <= r * r) break;
            if (all) break;
            if (full) {
                r = fmaf(sqrtf(kd), 1e-6f, sqrtf(kd));
                thr_d = kd; thr_i = NONE;                           // restart: every key with d <= kd is admitted again
            } else {
                r = 2.f * r;
            }
        }
        const size_t o = (size_t)batch * P + __float_as_int(ap.w);
        if (lane == 0) { min_d[o] = ld; min_i[o] = li; }
        if (lane < k_out) knn[o * k_out + lane] = li;
    }
}

// sum of a float array into a double accumulator (atomic per block)
__global__ void __launch_bounds__(256) k_sum(const float* __restrict__ x, long long n, double* __restrict__ out) {
    __shared__ double sd[33];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        acc += (double)x[i];
    acc = block_sum<double>(acc, sd);
    if (threadIdx.x == 0) atomicAdd(out, acc);
}

__global__ void k_finalize(const double* __restrict__ acc, int n, double scale, float* __restrict__ out) {
    const int i = threadIdx.x;
    if (i < n) out[i] = (float)(acc[i] * scale);
}

// d/da of  sum_i |a_i - b_{idx_a[i]}|^2  (scaled by *g_a)  and of  sum_j |a_{idx_b[j]} - b_j|^2  (scaled by *g_b)
__global__ void __launch_bounds__(256) k_chamfer_bwd(const float* __restrict__ a, const float* __restrict__ b, int P, int Q,
                                                     const int32_t* __restrict__ idx_a, const int32_t* __restrict__ idx_b,
                                                     const float* __restrict__ g_a, const float* __restrict__ g_b,
                                                     float scale, float* __restrict__ ga, float* __restrict__ gb) {
    const int batch = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const float ca = 2.f * scale * (*g_a), cb = 2.f * scale * (*g_b);
    if (t < P) {
        const size_t i = (size_t)batch * P + t, j = (size_t)batch * Q + idx_a[i];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float diff = ca * (a[3 * i + d] - b[3 * j + d]);
            if (ga) atomicAdd(ga + 3 * i + d, diff);
            if (gb) atomicAdd(gb + 3 * j + d, -diff);
        }
    }
    if (t < Q) {
        const size_t j = (size_t)batch * Q + t, i = (size_t)batch * P + idx_b[j];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float diff = cb * (a[3 * i + d] - b[3 * j + d]);
            if (ga) atomicAdd(ga + 3 * i + d, diff);
            if (gb) atomicAdd(gb + 3 * j + d, -diff);
        }
    }
}

template <int K>
static void launch_nn(const float4* a, const float4* b, const float2* trange_b, int B, int P, int Q, int prune, float* min_d,
                      int32_t* min_i, int32_t* knn, int k, cudaStream_t s) {
    k_nn<K><<<dim3(ceil_div(P, THREADS * QPT), B), THREADS, 0, s>>>(a, b, trange_b, P, Q, prune, min_d, min_i, knn, k);
}

// packs (and x-bucket-sorts, when the cloud fits the shared-memory pass) B clouds of P points into float4 records
static bool pack_cloud(const float* pts, int B, int P, float4* out, float2* trange, cudaStream_t s) {
    if (P <= SORT_MAX) {
        const size_t smem = (size_t)NBUCKET * 4 + (size_t)P * 4;
        static SmemOptIn optin;
        ensure_dynamic_smem(k_sort_x, NBUCKET * 4 + SORT_MAX * 4, optin, "knn_fwd (x sort)");   // failure surfaces at the launch check
        k_sort_x<<<B, 1024, smem, s>>>(pts, P, out, trange, ceil_div(P, TILE));
        return true;
    }
    const long long n = (long long)B * P;
    k_pack<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(pts, n, P, out);
    return false;
}

}  // namespace chamfer
}  // namespace mrb

using namespace mrb;
using namespace mrb::chamfer;

namespace {
struct Workspace {
    float4 *pa, *pb;
    float2 *tra, *trb;
    GridHdr *ha, *hb;
    uint16_t *csa, *csb;
    long long bytes;
};
// carves the caller's workspace; with base == 0 only the size is computed
Workspace carve(uintptr_t base, int B, int P, int Q) {
    Workspace w;
    uintptr_t p = (base + 15) & ~(uintptr_t)15;
    w.pa = reinterpret_cast<float4*>(p); p += sizeof(float4) * (size_t)B * P;
    w.pb = reinterpret_cast<float4*>(p); p += sizeof(float4) * (size_t)B * Q;
    w.tra = reinterpret_cast<float2*>(p); p += sizeof(float2) * (size_t)B * ceil_div(max(P, 1), TILE);
    w.trb = reinterpret_cast<float2*>(p); p += sizeof(float2) * (size_t)B * ceil_div(max(Q, 1), TILE);
    p = (p + 15) & ~(uintptr_t)15;
    w.ha = reinterpret_cast<GridHdr*>(p); p += sizeof(GridHdr) * (size_t)B;
    w.hb = reinterpret_cast<GridHdr*>(p); p += sizeof(GridHdr) * (size_t)B;
    w.csa = reinterpret_cast<uint16_t*>(p); p += sizeof(uint16_t) * (size_t)B * CS_STRIDE;
    w.csb = reinterpret_cast<uint16_t*>(p); p += sizeof(uint16_t) * (size_t)B * CS_STRIDE;
    w.bytes = (long long)(p - base) + 16;
    return w;
}
}  // namespace

extern "C" long long mrb_knn_workspace_bytes(int B, int P, int Q) {
    if (B < 0 || P < 0 || Q < 0) return -1;
    return carve(0, B, P, Q).bytes;
}

// algo: 0 = automatic (cell grid when both clouds have <= 65535 points, else the tiled scan), 1 = tiled scan, 2 = cell grid
extern "C" int mrb_knn_fwd_algo(const float* a, const float* b, int B, int P, int Q, int k, float* min_d_a, int32_t* min_i_a,
                                int32_t* knn_a, float* min_d_b, int32_t* min_i_b, int32_t* knn_b, void* workspace, int algo,
                                void* stream_) {
    MRB_REQUIRE(a && b && workspace, "knn_fwd: null pointer");
    MRB_REQUIRE(k >= 0 && k <= 16, "knn_fwd: k must be in [0, 16], got %d", k);
    MRB_REQUIRE((min_d_a && min_i_a) || (min_d_b && min_i_b), "knn_fwd: no output requested");
    MRB_REQUIRE(k == 0 || ((!min_d_a || knn_a) && (!min_d_b || knn_b)), "knn_fwd: knn output missing");
    MRB_REQUIRE(k <= Q && k <= P, "knn_fwd: k = %d exceeds a cloud size (%d, %d)", k, P, Q);
    MRB_REQUIRE(B == 0 || (P > 0 && Q > 0), "knn_fwd: empty cloud");
    MRB_REQUIRE(B <= 65535, "knn_fwd: batch too large");
    MRB_REQUIRE(algo >= 0 && algo <= 2, "knn_fwd: unknown algo %d", algo);
    const bool grid_ok = P <= GRID_MAX_POINTS && Q <= GRID_MAX_POINTS;
    MRB_REQUIRE(algo != 2 || grid_ok, "knn_fwd: the cell-grid path holds at most %d points per cloud", GRID_MAX_POINTS);
    if (B == 0) return MRB_OK;
    cudaStream_t s = (cudaStream_t)stream_;
    const Workspace w = carve((uintptr_t)workspace, B, P, Q);
    if (algo == 2 || (algo == 0 && grid_ok)) {
        const int Ga = grid_cells(P), Gb = grid_cells(Q);
        const int G = max(Ga, Gb);
        static SmemOptIn optin;
        if (int rc = ensure_dynamic_smem(k_grid_build, GMAX * GMAX * GMAX * 4, optin, "knn_fwd (cell grid)")) return rc;
        k_grid_build<<<2 * B, 1024, (size_t)G * G * G * 4, s>>>(a, b, B, P, Q, Ga, Gb, w.pa, w.pb, w.csa, w.csb, w.ha, w.hb,
                                                                sqrtf(GRID_C0 * (float)max(k, 1)));
#define MRB_NNG(KK)                                                                                                   \
    do {                                                                                                              \
        if (min_d_a)                                                                                                  \
            k_nn_grid<KK><<<dim3(ceil_div(P, (GRID_THREADS / 32) * GRID_QPW), B), GRID_THREADS, 0, s>>>(w.pa, w.pb, w.csb, w.hb, P, Q, min_d_a, \
                                                                                      min_i_a, knn_a, k);             \
        if (min_d_b)                                                                                                  \
            k_nn_grid<KK><<<dim3(ceil_div(Q, (GRID_THREADS / 32) * GRID_QPW), B), GRID_THREADS, 0, s>>>(w.pb, w.pa, w.csa, w.ha, Q, P, min_d_b, \
                                                                                      min_i_b, knn_b, k);             \
    } while (0)
        if (k <= 1) MRB_NNG(1);
        else if (k <= 4) MRB_NNG(4);
        else if (k <= 10) MRB_NNG(10);
        else MRB_NNG(16);
#undef MRB_NNG
        return check_launch("knn_fwd (grid)");
    }
    const bool sa = pack_cloud(a, B, P, w.pa, w.tra, s);
    const bool sb = pack_cloud(b, B, Q, w.pb, w.trb, s);
#define MRB_NN(KK)                                                                                        \
    do {                                                                                                  \
        if (min_d_a) launch_nn<KK>(w.pa, w.pb, w.trb, B, P, Q, sb ? 1 : 0, min_d_a, min_i_a, knn_a, k, s);  \
        if (min_d_b) launch_nn<KK>(w.pb, w.pa, w.tra, B, Q, P, sa ? 1 : 0, min_d_b, min_i_b, knn_b, k, s);  \
    } while (0)
    if (k <= 1) MRB_NN(1);
    else if (k <= 4) MRB_NN(4);
    else if (k <= 10) MRB_NN(10);
    else MRB_NN(16);
#undef MRB_NN
    return check_launch("knn_fwd");
}

extern "C" int mrb_knn_fwd(const float* a, const float* b, int B, int P, int Q, int k, float* min_d_a, int32_t* min_i_a,
                           int32_t* knn_a, float* min_d_b, int32_t* min_i_b, int32_t* knn_b, void* workspace,
                           void* stream_) {
    return mrb_knn_fwd_algo(a, b, B, P, Q, k, min_d_a, min_i_a, knn_a, min_d_b, min_i_b, knn_b, workspace, 0, stream_);
}

extern "C" int mrb_sum_scaled(const float* x, long long n, double scale, double* acc, float* out, void* stream_) {
    MRB_REQUIRE(x && acc && out, "sum_scaled: null pointer");
    cudaStream_t s = (cudaStream_t)stream_;
    cudaMemsetAsync(acc, 0, sizeof(double), s);
    if (n > 0) {
        const int blocks = (int)min((long long)4 * kNumSMs, ceil_div64(n, 256));
        k_sum<<<blocks, 256, 0, s>>>(x, n, acc);
    }
    k_finalize<<<1, 32, 0, s>>>(acc, 1, scale, out);
    return check_launch("sum_scaled");
}

extern "C" int mrb_chamfer_bwd(const float* a, const float* b, int B, int P, int Q, const int32_t* idx_a,
                               const int32_t* idx_b, const float* g_a, const float* g_b, float scale, float* ga,
                               float* gb, void* stream_) {
    MRB_REQUIRE(a && b && idx_a && idx_b && g_a && g_b, "chamfer_bwd: null pointer");
    if (B == 0 || (P == 0 && Q == 0) || (!ga && !gb)) return MRB_OK;
    k_chamfer_bwd<<<dim3(ceil_div(max(P, Q), 256), B), 256, 0, (cudaStream_t)stream_>>>(a, b, P, Q, idx_a, idx_b, g_a,
                                                                                       g_b, scale, ga, gb);
    return check_launch("chamfer_bwd");
}

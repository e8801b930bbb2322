// GraphConv / linear projections on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
//   C[M x N] = A[M x K] * B[K x N]      A: fp32 activations, row-major (lda), M = number of vertices (50k .. 7M)
//                                       B: weights, pre-packed once per call into a K-major, 128B-swizzled, split image
//
// Replaces the cuBLAS sgemm call sites of the reference (meshRCNN/layers.py:54,57 torch.mm(x, w0/w1), nn.Linear at
// :93,155,230,255,335) and their autograd transposes.  fp32 parity (rtol 1e-4) rules out a single TF32/BF16 pass, so
// the product is evaluated as a 3xTF32 split on the tensor pipe:  a = a_hi + a_lo (both tf32-representable),
//   a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi          (error ~2^-21 relative, fp32 accumulate in TMEM)
// Arithmetic intensity at N <= 256 is below the B200 ridge, so the kernel is HBM-bound by design: A is read once,
// C written once, and the weight image (<= 0.8 MB) streams from L2.
//
// Persistent CTAs (one per SM) loop over 128-row M tiles x N tiles (<= 256 columns).  Warp roles (576 threads):
//   warps 0-7   A producers: coalesced fp32 loads of the A chunk (128 x 32), 3 chunks in flight per thread; the hi/lo split
//               and the next loads are issued BEFORE the wait for the stage, after it only swizzled st.shared,
//               fence.proxy.async, mbarrier arrive
//   warps 8-15  epilogue: tcgen05.ld (TMEM lane quadrant = warp % 4, column half = warp / 12) -> SWIZZLE_128B shared-memory
//               tile -> TMA tensor store of the 32 x 32 block (cp.async.bulk.tensor, or cp.reduce...add for K segments) when
//               C rows are 16-byte aligned, guarded scalar stores otherwise; the TMEM accumulator is double buffered, so
//               tile j's write-back overlaps tile j+1's MMAs
//   warp 16     lane 0 issues tcgen05.mma (kind::tf32, M=128, N=NT, K=8), tcgen05.commit frees the stage / signals the
//               epilogue; the whole warp owns the TMEM allocation
//   warp 17     lane 0 streams the packed weight chunk with cp.async.bulk (TMA bulk copy) onto the stage's mbarrier
#include <cuda.h>
#include <string.h>
#include "common.cuh"
#include "../../include/meshrcnn_b200.h"

namespace mrb {
namespace gemmtc {

constexpr int BM = 128;            // rows per CTA (UMMA M)
constexpr int BK = 32;             // fp32 elements per K chunk = one 128-byte swizzle row
constexpr int STAGES = 2;
constexpr int NT_MAX = 256;        // max columns per CTA (UMMA N)
constexpr int A_BYTES = BM * BK * 4;                   // 16 KB per (hi | lo) tile
constexpr int B_BYTES_MAX = NT_MAX * BK * 4;           // 32 KB per (hi | lo) tile
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES_MAX;   // 96 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int THREADS = 192;

// byte offset of element (row, k) inside a K-major SWIZZLE_128B tile whose rows are 32 fp32 wide
__host__ __device__ __forceinline__ int swz(int row, int k) {
    return (row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 2) ^ (row & 7)) & 7) << 4) + (k & 3) * 4;
}

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(=1)<<16 |
// SBO(=1024>>4)<<32 | version(1)<<46 | layout SWIZZLE_128B(2)<<61
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), K-major both, N>>3 @17, M>>4 @24
__device__ __forceinline__ uint32_t umma_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns of TMEM -> 32 registers per thread (thread = lane / row, register = column)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
          "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
          "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

struct Params {
    const float* A;
    int lda, M, K;
    const unsigned char* image;   // [n_tile][chunk][hi|lo][NT rows][128 B]
    int N, NT, nchunks, tmem_cols;   // tmem_cols: columns of one accumulator
    int nacc;                        // accumulators used round-robin over the K chunks (bounds the fp32 chain length)
    int acc_bufs;                    // 2: the epilogue of tile j overlaps the MMAs of tile j+1 (double-buffered TMEM)
    int mtiles, ntiles;
    int accumulate;                  // epilogue adds to C (K is processed in segments, see mrb_gemm_tc)
    int stages, stage_bytes;         // shared-memory ring: stages x (A hi | A lo | B hi | B lo)
    long long img_tile_stride;       // bytes between the images of consecutive N tiles
    float* C;
    int ldc;
    int tma_store;                   // cmap describes C: the epilogue writes 32 x 32 blocks with TMA tensor stores
};

constexpr int PROD_WARPS = 8;        // A producer warps, 16 rows of the 128-row tile each
constexpr int PROD_ROWS = BM / PROD_WARPS;
// (measured with the TMA-store epilogue, K=128->256 / 256->128 / 256->256 at 206 k rows: 1: 100.9 / 141.9 / 164.0 us, 2: 100.0 /
//  142.2 / 161.7, 3: 90.4 / 131.9 / 146.1, 4 (spills): 107.4 / 163.3 / 182.6)
#ifndef MRB_TC_PREFETCH
#define MRB_TC_PREFETCH 3
#endif
constexpr int PREFETCH = MRB_TC_PREFETCH;   // A chunks in flight per producer thread (registers)
// Diagnostic builds (scripts/variants.sh gemm_tc.cu ...): -DMRB_DIAG_NOLOAD (producers skip the global loads), -DMRB_DIAG_NOMMA
// (the issuer skips the MMAs), -DMRB_DIAG_NOSTORE (the epilogue skips the C stores), -DMRB_DIAG_NOEPI (the epilogue stops after
// its TMEM reads), -DMRB_DIAG_HALFB (half the weight bytes per chunk) isolate the legs of the pipeline (timing only);
// -DMRB_DIAG_TN_NOLOAD / _NOEPI / _NOMMA do the same for the weight-gradient kernel.  Results: profiles/r02/gemm_tc_pipeline_legs.log.
constexpr int EPI_WARPS = 8;         // epilogue warps: TMEM lane quadrant = warp % 4, column half = (warp - PROD_WARPS) / 4
constexpr int TC_THREADS = (PROD_WARPS + EPI_WARPS + 2) * 32;   // producers | epilogue warps | MMA issuer | weight TMA
constexpr int EPI_BYTES = EPI_WARPS * 32 * 33 * 4;
constexpr int TC_MAX_STAGES = 4;
#ifndef MRB_TC_SINGLE_ACC_CHUNKS
#define MRB_TC_SINGLE_ACC_CHUNKS 16
#endif
constexpr int TC_SINGLE_ACC_CHUNKS = MRB_TC_SINGLE_ACC_CHUNKS;   // K <= 512: one TMEM accumulator for all three 3xTF32 products
constexpr int TC_SMEM_LIMIT = 232448;                                    // 227 KB opt-in maximum per CTA
constexpr int TC_RING_BYTES = TC_SMEM_LIMIT - EPI_BYTES - 1024 - 256;   // what the stage ring may use

#ifdef MRB_TC_TIMELINE
// Diagnostic build only (scripts/gemm_timeline.sh): per-role clock64() stamps of CTA 0.
__device__ long long g_tl[4][2048];
__device__ int g_tl_n[4];
#define TL(role, tag) do { if (blockIdx.x == 0 && tl_n < 2048) { g_tl[role][tl_n++] = (clock64() << 8) | (tag); g_tl_n[role] = tl_n; } } while (0)
#else
#define TL(role, tag) do { } while (0)
#endif

// Persistent: grid = min(#tiles, #SMs); every role loops over the CTA's tiles with pipeline state that carries across
// tiles, so global-load latency, tensor work and the C write-back of consecutive tiles overlap on one SM.
__global__ void __launch_bounds__(TC_THREADS, 1) k_gemm_tc(Params p, const __grid_constant__ CUtensorMap cmap) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int STAGES = p.stages, STAGE_BYTES = p.stage_bytes;
    float* epi = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);   // full[4] empty[4] tfull[2] tempty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (TC_MAX_STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * TC_MAX_STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * TC_MAX_STAGES + 2 + a); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef MRB_TC_TIMELINE
    int tl_n = 0;
#endif
    const int NT = p.NT;
    const int b_bytes = NT * BK * 4;     // one (hi | lo) weight tile
    const int total_tiles = p.mtiles * p.ntiles;
    const int my_tiles = (total_tiles > (int)blockIdx.x) ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int acc_cols = p.tmem_cols * p.nacc;
    const int tmem_total = acc_cols * p.acc_bufs;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), PROD_WARPS + 1);     // one elected arrive per producer warp + the weight TMA's expect_tx
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == PROD_WARPS + EPI_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(tmem_total)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < PROD_WARPS) {
        // ===== A producers ======================================================================================
        // The kernel is bound by global-load latency unless enough bytes are in flight, so every producer thread keeps
        // PREFETCH chunks (16 coalesced 128-byte row segments each) outstanding in registers, across tile boundaries.
        const int r0 = warp * PROD_ROWS;
        const int total_its = my_tiles * p.nchunks;
        if (((p.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0)) {
            // 16-byte aligned rows: lane = (row r0 + 4 i + lane / 8, 16-byte k chunk lane % 8) -> 4 x ld.global.v4 and
            // 8 x st.shared.v4 per chunk instead of 16 + 32 scalar accesses (4x fewer requests in the L1 miss queue, which
            // is what bounds the scalar path); a quarter-warp writes one complete swizzled 128-byte row (conflict free).
            const int rsub = lane >> 3, kc = lane & 7;
            float4 v[PREFETCH][PROD_ROWS / 4];
            auto load_it = [&](int it, float4 (&dst)[PROD_ROWS / 4]) {
                const int tile = blockIdx.x + (it / p.nchunks) * gridDim.x;
                const int c = it % p.nchunks;
                const int m0 = (tile / p.ntiles) * BM;
                // Always-valid (clamped) addresses and NO use of the loaded value here: masking happens at store time, so the
                // loads of PREFETCH chunks really stay in flight (a select on the loaded value would wait for it right away).
                const int k = min(c * BK + kc * 4, p.lda - 4);
#pragma unroll
                for (int i = 0; i < PROD_ROWS / 4; ++i) {
                    const int gm = min(m0 + r0 + 4 * i + rsub, p.M - 1);
#ifdef MRB_DIAG_NOLOAD
                    dst[i] = make_float4((float)gm, (float)k, 1.f, 2.f);
#else
                    dst[i] = __ldg(reinterpret_cast<const float4*>(p.A + (size_t)gm * p.lda + k));
#endif
                }
            };
#pragma unroll
            for (int d = 0; d < PREFETCH; ++d)
                if (d < total_its) load_it(d, v[d]);
            for (int it0 = 0; it0 < total_its; it0 += PREFETCH) {
#pragma unroll
                for (int d = 0; d < PREFETCH; ++d) {
                    const int it = it0 + d;
                    if (it < total_its) {
                        // Split into tf32 hi / lo BEFORE waiting for the stage (the values sit in registers anyway) and re-issue the
                        // registers' next global loads right away: after the wait only the 8 shared-memory stores remain on the
                        // stage's turn-around path.
                        const int s = it % STAGES, use = it / STAGES;
                        const int m0 = ((blockIdx.x + (it / p.nchunks) * gridDim.x) / p.ntiles) * BM;
                        const int kleft = p.K - ((it % p.nchunks) * BK + kc * 4);   // valid columns of this 16-byte chunk
                        float4 hi[PROD_ROWS / 4], lo[PROD_ROWS / 4];
#pragma unroll
                        for (int i = 0; i < PROD_ROWS / 4; ++i) {
                            float4 x = v[d][i];
                            const bool rin = m0 + r0 + 4 * i + rsub < p.M;
                            x.x = (rin && kleft > 0) ? x.x : 0.f;       // columns [K, lda) of a row are padding (may hold anything)
                            x.y = (rin && kleft > 1) ? x.y : 0.f;
                            x.z = (rin && kleft > 2) ? x.z : 0.f;
                            x.w = (rin && kleft > 3) ? x.w : 0.f;
                            hi[i].x = tf32_rna(x.x); hi[i].y = tf32_rna(x.y); hi[i].z = tf32_rna(x.z); hi[i].w = tf32_rna(x.w);
                            lo[i].x = tf32_rna(x.x - hi[i].x); lo[i].y = tf32_rna(x.y - hi[i].y); lo[i].z = tf32_rna(x.z - hi[i].z);
                            lo[i].w = tf32_rna(x.w - hi[i].w);
                        }
                        if (it + PREFETCH < total_its) load_it(it + PREFETCH, v[d]);
                        if (use > 0) mbar_wait(empty_bar(s), (use - 1) & 1);
                        if (threadIdx.x == 0) TL(0, 50);
                        unsigned char* a_hi = smem + s * STAGE_BYTES;
                        unsigned char* a_lo = a_hi + A_BYTES;
#pragma unroll
                        for (int i = 0; i < PROD_ROWS / 4; ++i) {
                            const int row = r0 + 4 * i + rsub;
                            const int off = (row >> 3) * 1024 + (row & 7) * 128 + (((kc ^ (row & 7)) & 7) << 4);
                            *reinterpret_cast<float4*>(a_hi + off) = hi[i];
                            *reinterpret_cast<float4*>(a_lo + off) = lo[i];
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(full_bar(s));
                        if (threadIdx.x == 0) TL(0, 51);
                    }
                }
            }
        } else {
            float v[PREFETCH][PROD_ROWS];
            auto load_it = [&](int it, float (&dst)[PROD_ROWS]) {
                const int tile = blockIdx.x + (it / p.nchunks) * gridDim.x;
                const int c = it % p.nchunks;
                const int m0 = (tile / p.ntiles) * BM;
                const int k = min(c * BK + lane, p.K - 1);      // clamped, masked at store time (see the vector path)
#pragma unroll
                for (int i = 0; i < PROD_ROWS; ++i) {
                    const int gm = min(m0 + r0 + i, p.M - 1);
                    dst[i] = __ldg(p.A + (size_t)gm * p.lda + k);
                }
            };
#pragma unroll
            for (int d = 0; d < PREFETCH; ++d)
                if (d < total_its) load_it(d, v[d]);
            for (int it0 = 0; it0 < total_its; it0 += PREFETCH) {
#pragma unroll
                for (int d = 0; d < PREFETCH; ++d) {
                    const int it = it0 + d;
                    if (it < total_its) {
                        const int s = it % STAGES, use = it / STAGES;
                        if (use > 0) mbar_wait(empty_bar(s), (use - 1) & 1);
                        unsigned char* a_hi = smem + s * STAGE_BYTES;
                        unsigned char* a_lo = a_hi + A_BYTES;
                        const int m0 = ((blockIdx.x + (it / p.nchunks) * gridDim.x) / p.ntiles) * BM;
                        const bool kin = (it % p.nchunks) * BK + lane < p.K;
#pragma unroll
                        for (int i = 0; i < PROD_ROWS; ++i) {
                            const float x = (kin && m0 + r0 + i < p.M) ? v[d][i] : 0.f;
                            const float hi = tf32_rna(x);
                            const float lo = tf32_rna(x - hi);
                            const int off = swz(r0 + i, lane);
                            *reinterpret_cast<float*>(a_hi + off) = hi;
                            *reinterpret_cast<float*>(a_lo + off) = lo;
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA
                        __syncwarp();
                        if (lane == 0) mbar_arrive(full_bar(s));    // 256 per-thread arrives on one mbarrier serialise (~10 cycles each)
                        if (it + PREFETCH < total_its) load_it(it + PREFETCH, v[d]);
                    }
                }
            }
        }
    } else if (warp < PROD_WARPS + EPI_WARPS) {
        // ===== epilogue: TMEM -> registers -> C =========================================================================
        // Warp w reads TMEM lane quadrant w % 4 (= 32 rows of the tile) and the column blocks of its half.  16-byte aligned
        // C rows are written straight from registers (each thread owns 32 consecutive columns of one row -> 8 x st.v4);
        // otherwise the block is transposed through shared memory so that every store instruction covers one 128-byte row
        // segment.
        const int ew = warp & 3;                       // TMEM lane quadrant (PROD_WARPS is a multiple of 4)
        const int half = (warp - PROD_WARPS) >> 2;     // 0 | 1
        float* stage_t = epi + (warp - PROD_WARPS) * (32 * 33);
        float* stage_tma = epi + (warp - PROD_WARPS) * (32 * 32);      // 4096-byte tiles (the ring keeps epi 1024-byte aligned)
        const bool aligned = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
        const int nblk = (NT + 31) / 32;
        const int blk_beg = half * ((nblk + 1) / 2), blk_end = half ? nblk : (nblk + 1) / 2;
        int j = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++j) {
            const int ab = j % p.acc_bufs, use = j / p.acc_bufs;
            const int m0 = (tile / p.ntiles) * BM, n0 = (tile % p.ntiles) * NT;
            mbar_wait(tfull_bar(ab), use & 1);
            if (warp == PROD_WARPS && lane == 0) TL(1, 40);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + (uint32_t)(ab * acc_cols);
            const int gm_mine = m0 + ew * 32 + lane;
            for (int blk = blk_beg; blk < blk_end; ++blk) {
                const int c0 = blk * 32;
                float acc[32];
                {
                    // both accumulator reads are in flight before the single wait
                    uint32_t v0[32], v1[32];
                    const uint32_t taddr = tmem_d + ((uint32_t)(ew * 32) << 16) + (uint32_t)c0;
                    tmem_ld32(taddr, v0);
                    if (p.nacc > 1) tmem_ld32(taddr + (uint32_t)p.tmem_cols, v1);   // main + cross-term accumulator
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (warp == PROD_WARPS && lane == 0) TL(1, 43);
                    if (p.nacc > 1) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) acc[q] = __uint_as_float(v0[q]) + __uint_as_float(v1[q]);
                    } else {
#pragma unroll
                        for (int q = 0; q < 32; ++q) acc[q] = __uint_as_float(v0[q]);
                    }
                    for (int a = 2; a < p.nacc; ++a) {
                        tmem_ld32(taddr + (uint32_t)(a * p.tmem_cols), v1);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int q = 0; q < 32; ++q) acc[q] += __uint_as_float(v1[q]);
                    }
                }
                if (blk + 1 == blk_end) {
                    // last read of this accumulator by this warp: hand the TMEM buffer back before the stores
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(ab));
                    if (warp == PROD_WARPS && lane == 0) TL(1, 41);
                }
                const int colbase = n0 + c0;
#ifdef MRB_DIAG_NOEPI      // timing experiment only: the epilogue ends after the TMEM reads (no transpose, no stores)
                if (acc[0] != 1.2345e-33f) continue;
#endif
                if (p.tma_store) {
                    // Each thread holds 32 columns of ONE row.  The block is written into a SWIZZLE_128B shared-memory tile
                    // (16-byte chunk q of row r at chunk q ^ (r & 7): conflict free) and handed to the TMA unit, which writes
                    // the 32 x 32 block to C (or adds it, for K segments after the first) and clips the rows beyond M; the warp
                    // only waits until the tile has been READ before it refills it.
                    float4* st4 = reinterpret_cast<float4*>(stage_tma);        // 32 rows x 8 chunks of 16 bytes, 1024-byte aligned
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        st4[lane * 8 + (q ^ (lane & 7))] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
                    if (c0 + 32 <= NT && colbase + 32 <= p.N) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (warp == PROD_WARPS && lane == 0) TL(1, 44);
#ifndef MRB_DIAG_NOSTORE
                        if (lane == 0) {
                            const uint32_t src = smem_u32(stage_tma);
                            const int row0 = m0 + ew * 32;
                            if (p.accumulate)
                                asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&cmap),
                                             "r"(src), "r"(colbase), "r"(row0)
                                             : "memory");
                            else
                                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&cmap),
                                             "r"(src), "r"(colbase), "r"(row0)
                                             : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
#endif
                    } else {
                        // trailing partial column block (N % 32 != 0): the tensor store clips only at 16-byte granularity, so the
                        // block is written from the same tile with guarded, coalesced scalar stores (lane = column)
                        __syncwarp();
                        const int col = colbase + lane;
                        if (c0 + lane < NT && col < p.N) {
                            const float* tile = stage_tma;
                            const int rmax = min(32, p.M - (m0 + ew * 32));
                            float* dst = p.C + (size_t)(m0 + ew * 32) * p.ldc + col;
                            for (int r = 0; r < rmax; ++r) {
                                const float v = tile[r * 32 + ((((lane >> 2) ^ (r & 7)) << 2) | (lane & 3))];
                                dst[(size_t)r * p.ldc] = v + (p.accumulate ? dst[(size_t)r * p.ldc] : 0.f);
                            }
                        }
                    }
                    if (warp == PROD_WARPS && lane == 0) TL(1, 45);
                } else if (aligned && c0 + 32 <= NT && colbase + 32 <= p.N) {
                    // Each thread holds 32 columns of ONE row; storing them directly would touch 32 different 128-byte lines
                    // per instruction.  The block goes through a swizzled (conflict-free) shared-memory tile instead, so
                    // that every st.global.v4 instruction writes four complete 128-byte row segments.
                    float4* st4 = reinterpret_cast<float4*>(stage_t);          // 32 rows x 8 chunks of 16 bytes
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        st4[lane * 8 + (q ^ (lane & 7))] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
                    __syncwarp();
                    if (warp == PROD_WARPS && lane == 0) TL(1, 44);
                    const int rsub = lane >> 3, ch = lane & 7;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int r = 4 * j + rsub;
                        float4 o = st4[r * 8 + (ch ^ (r & 7))];
                        const int gm = m0 + ew * 32 + r;
                        if (gm < p.M) {
                            float4* dst = reinterpret_cast<float4*>(p.C + (size_t)gm * p.ldc + colbase) + ch;
                            if (p.accumulate) {
                                const float4 old = *dst;
                                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                            }
#ifdef MRB_DIAG_NOSTORE
                            if (o.x == 1.2345e-33f)
#endif
                            *dst = o;
                        }
                    }
                    if (warp == PROD_WARPS && lane == 0) TL(1, 45);
                } else {
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 32; ++q) stage_t[lane * 33 + q] = acc[q];
                    __syncwarp();
                    const int col = colbase + lane;
                    if (c0 + lane < NT && col < p.N) {
                        float* dst = p.C + (size_t)(m0 + ew * 32) * p.ldc + col;
                        const int rmax = min(32, p.M - (m0 + ew * 32));
#pragma unroll 8
                        for (int r = 0; r < 32; ++r)
                            if (r < rmax) dst[(size_t)r * p.ldc] = stage_t[r * 33 + lane] + (p.accumulate ? dst[(size_t)r * p.ldc] : 0.f);
                    }
                    __syncwarp();
                }
            }
            if (warp == PROD_WARPS && lane == 0) TL(1, 42);
            if (blk_beg >= blk_end) {                 // this half has no column block (NT <= 32): still release the buffer
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(ab));
            }
        }
        if (p.tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all tensor stores of this warp done
    } else if (warp == PROD_WARPS + EPI_WARPS) {
        // ===== MMA issuer ========================================================================================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(NT);
            int it = 0, j = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++j) {
                const int ab = j % p.acc_bufs, use = j / p.acc_bufs;
                if (use > 0) {
                    mbar_wait(tempty_bar(ab), (use - 1) & 1);          // epilogue has drained this accumulator
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint32_t tmem_d = tmem_base + (uint32_t)(ab * acc_cols);
                TL(2, 10);
                for (int c = 0; c < p.nchunks; ++c, ++it) {
                    const int s = it % STAGES, suse = it / STAGES;
                    mbar_wait(full_bar(s), suse & 1);
                    TL(2, 20);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_base + s * STAGE_BYTES, a_lo = a_hi + A_BYTES;
                    const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + b_bytes;
                    const int ksteps = min(BK / 8, (p.K - c * BK + 7) / 8);     // the last chunk may be mostly padding
#pragma unroll
                    for (int kk = 0; kk < BK / 8; ++kk) {
                        if (kk >= ksteps) break;
                        const uint64_t dah = umma_desc(a_hi + kk * 32), dal = umma_desc(a_lo + kk * 32);
                        const uint64_t dbh = umma_desc(b_hi + kk * 32), dbl = umma_desc(b_lo + kk * 32);
                        // The fp32 accumulate of the tensor core truncates (error ~1 ulp of the accumulator per MMA, biased), so
                        // the two small cross terms get their own accumulator (2^-11 of the magnitude: their truncation is
                        // negligible) and only the K/8 hi*hi products are chained into the main one(s).
                        // (short reductions -- nacc == 1 -- chain all three products into one accumulator, like the weight-
                        // gradient kernel does for <= 384 MMAs: that leaves room for a second accumulator set, so the epilogue of
                        // a 256-column tile overlaps the next tile's MMAs)
                        const int nmain = max(p.nacc - 1, 1);
                        const uint32_t d_main = tmem_d + (uint32_t)((c % nmain) * p.tmem_cols);
                        const uint32_t d_aux = p.nacc > 1 ? tmem_d + (uint32_t)(nmain * p.tmem_cols) : d_main;
#ifndef MRB_DIAG_NOMMA
                        umma_tf32(d_aux, dal, dbh, idesc, (c | kk) != 0);
                        umma_tf32(d_aux, dah, dbl, idesc, 1);
                        umma_tf32(d_main, dah, dbh, idesc, p.nacc > 1 ? ((c >= nmain) || kk != 0) : 1);
#endif
                    }
                    umma_commit(empty_bar(s));                 // stage reusable once these MMAs retire
                    TL(2, 21);
                }
                umma_commit(tfull_bar(ab));                    // accumulator complete -> epilogue
            }
        }
        __syncwarp();
    } else {
        // ===== weight-chunk TMA (bulk copy) producer =======================================================================
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const unsigned char* img = p.image + (size_t)(tile % p.ntiles) * p.img_tile_stride;
                for (int c = 0; c < p.nchunks; ++c, ++it) {
                    const int s = it % STAGES, use = it / STAGES;
                    if (use > 0) mbar_wait(empty_bar(s), (use - 1) & 1);
                    TL(3, 60);
                    const uint32_t dst = smem_base + s * STAGE_BYTES + 2 * A_BYTES;
#ifdef MRB_DIAG_HALFB      // timing experiment only (wrong results): what would half the weight bytes per SM buy?
                    mbar_expect_tx(full_bar(s), b_bytes);
                    bulk_g2s(dst, img + (size_t)c * 2 * b_bytes, b_bytes / 2, full_bar(s));
                    bulk_g2s(dst + b_bytes, img + (size_t)c * 2 * b_bytes + b_bytes, b_bytes / 2, full_bar(s));
#else
                    mbar_expect_tx(full_bar(s), 2 * b_bytes);
                    bulk_g2s(dst, img + (size_t)c * 2 * b_bytes, b_bytes, full_bar(s));
                    bulk_g2s(dst + b_bytes, img + (size_t)c * 2 * b_bytes + b_bytes, b_bytes, full_bar(s));
#endif
                }
            }
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == PROD_WARPS + EPI_WARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_total) : "memory");
    }
}

// Weight image: element (k, n) of the logical K x N operand, read through generic strides from up to two sources
// (so that [W0 | W1] and its transpose never have to be concatenated in memory), split into tf32 hi / lo and written
// in the exact shared-memory layout of the main kernel (zero padded in K and N).
struct PackParams {
    const float* src0;
    const float* src1;
    long long sk, sn;     // strides (elements) of k and n in the sources
    int split_axis;       // 0: single source; 1: n >= split_at comes from src1 (n - split_at); 2: same along k
    int split_at;
    int K, N, NT, nchunks, ntiles;
    unsigned char* image;
};

__device__ __forceinline__ void pack_element(const PackParams& p, long long t) {
    const long long total = (long long)p.ntiles * p.nchunks * p.NT * BK;
    if (t >= total) return;
    const int kl = (int)(t % BK);
    const int nl = (int)((t / BK) % p.NT);
    const int c = (int)((t / ((long long)BK * p.NT)) % p.nchunks);
    const int tile = (int)(t / ((long long)BK * p.NT * p.nchunks));
    const int k = c * BK + kl, n = tile * p.NT + nl;
    float v = 0.f;
    if (k < p.K && n < p.N) {
        const float* src = p.src0;
        int kk = k, nn = n;
        if (p.split_axis == 1 && n >= p.split_at) { src = p.src1; nn = n - p.split_at; }
        if (p.split_axis == 2 && k >= p.split_at) { src = p.src1; kk = k - p.split_at; }
        v = src[kk * p.sk + nn * p.sn];
    }
    const float hi = tf32_rna(v), lo = tf32_rna(v - hi);
    const size_t tile_bytes = (size_t)p.NT * BK * 4;
    unsigned char* base = p.image + ((size_t)tile * p.nchunks + c) * 2 * tile_bytes;
    const int off = swz(nl, kl);
    *reinterpret_cast<float*>(base + off) = hi;
    *reinterpret_cast<float*>(base + tile_bytes + off) = lo;
}

__global__ void k_pack_b(PackParams p) { pack_element(p, (long long)blockIdx.x * blockDim.x + threadIdx.x); }

// two images in one launch (a GraphConv's forward operand [W0 | W1] and the transposed operand of its input gradient)
__global__ void k_pack_b2(PackParams a, PackParams b, long long total_a) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total_a) pack_element(a, t);
    else pack_element(b, t - total_a);
}

// many images in one launch (all dense GraphConv blocks of a forward pass): block -> entry through a prefix table
constexpr int PACK_BATCH_MAX = 32;
struct PackBatch {
    PackParams e[PACK_BATCH_MAX];
    unsigned first_block[PACK_BATCH_MAX + 1];
    int n;
};
__global__ void __launch_bounds__(256) k_pack_batch(const __grid_constant__ PackBatch b) {
    int i = 0;
    while (i + 1 < b.n && blockIdx.x >= b.first_block[i + 1]) ++i;
    pack_element(b.e[i], (long long)(blockIdx.x - b.first_block[i]) * blockDim.x + threadIdx.x);
}

struct Plan {
    int NT, ntiles, nchunks, tmem_cols, nacc;
    int stages, stage_bytes;
    size_t image_bytes;
};

// Column tile: up to 256 columns per CTA tile.  (128-column tiles would let two accumulator sets fit the 512 TMEM columns so
// that the epilogue overlaps the next tile's MMAs, but every extra column tile re-produces the A rows; measured slower:
// 354 vs 239 us at M = 206k, K = 387, N = 256.)
static Plan make_plan(int K, int N) {
    Plan pl;
    pl.ntiles = (N + NT_MAX - 1) / NT_MAX;
    const int per = (N + pl.ntiles - 1) / pl.ntiles;
    pl.NT = ((per + 15) / 16) * 16;
    pl.nchunks = (K + BK - 1) / BK;
    int cols = 32;
    while (cols < pl.NT) cols <<= 1;
    pl.tmem_cols = cols;
    // The tensor-core fp32 accumulate truncates, so the error of one accumulator grows linearly with the number of MMAs
    // chained into it; long reductions (K > 512, e.g. the 3840 -> 128 bottleneck) are spread over up to 4 accumulators
    // (512 TMEM columns) that the epilogue adds on the CUDA cores.
    pl.nacc = 1;
    if (pl.nchunks > 16) {
        const int want = min(4, min(512 / cols, (pl.nchunks + 15) / 16));
        while (pl.nacc * 2 <= want) pl.nacc *= 2;      // TMEM allocations are powers of two
    }
    pl.image_bytes = (size_t)pl.ntiles * pl.nchunks * 2 * pl.NT * BK * 4;
    pl.stage_bytes = 2 * A_BYTES + 2 * pl.NT * BK * 4;
    pl.stages = min(TC_MAX_STAGES, TC_RING_BYTES / pl.stage_bytes);
    return pl;
}


// ==========================================================================================================
// Weight gradients:  C[Kin x N] += X^T[Kin x V] * G[V x N]   (reduction over the V vertices, split over CTAs)
// ==========================================================================================================
// In memory a vertex row of X / G is contiguous along M / N, i.e. the operands are "MN-major"; the producers transpose
// them on the fly into the same K-major SWIZZLE_128B tiles the forward kernel uses: lane = one row (i or j) of a 32-row
// block, four consecutive vertices form one 16-byte chunk -> one st.shared.v4 per (hi | lo), bank-conflict free because
// the swizzle spreads the 8 rows of a quarter-warp over the 8 chunks of a 128-byte line.  Global loads stay coalesced
// (a warp reads 32 consecutive floats of one vertex row).
// grid = (M tiles, splits); each CTA reduces its slice of vertices into TMEM and adds the tile to C with fp32
// reductions (C is zero-filled by the caller side of the C ABI).
struct ParamsTN {
    const float* X; int ldx;      // V x Kin
    const float* G; int ldg;      // V x N
    int V, Kin, N;                // N % 32 == 0, N <= 256
    int chunks_per_split;
    float* C0; float* C1;         // columns [0, n_split) -> C0 (ld ldc), [n_split, N) -> C1
    int n_split, ldc;
    int tail;                     // 0..TN_TAIL_MAX extra feature columns, accumulated on the CUDA cores by tile 0:
    const float* Xtail; int ld_tail;   //   T[m, j] += sum_v Xtail[v * ld_tail + m] * G[v, j]   (m < tail)
    float* T0; float* T1;              //   columns [0, n_split) -> T0 + m * ldc, [n_split, N) -> T1 + m * ldc
};
constexpr int TN_TAIL_MAX = 4;

constexpr int TN_PROD_WARPS = 8;                       // warp w stages vertices [4w, 4w+4) of every chunk = 16-byte k chunk w
constexpr int TN_THREADS = (TN_PROD_WARPS + 1) * 32;   // producers (also the epilogue) | MMA issuer
// NBMAX: 32-column blocks of G a producer thread stages (N <= 32 * NBMAX); PF: chunks in flight per producer thread.  The
// register file allows two chunks in flight only for N <= 128 (16 + 4 * NBMAX staging registers per chunk; 9 warps put 3 on
// one scheduler, which caps a thread at 168 registers: <8, 2> spills 656 bytes and runs 2x slower, 119 vs 58 us at V = 50k).
template <int NBMAX, int TN_PREFETCH>
__global__ void __launch_bounds__(TN_THREADS, 1) k_gemm_tn(ParamsTN p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    const uint32_t tfull_bar = bar_base + 8u * (2 * STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef MRB_TC_TIMELINE
    int tl_n = 0;
#define TLY(role, tag) do { if (blockIdx.y == 0) TL(role, tag); } while (0)
#else
#define TLY(role, tag) do { } while (0)
#endif
    const int i0 = blockIdx.x * BM;
    const int NT = p.N;
    const int b_bytes = NT * BK * 4;
    const int total_chunks = (p.V + BK - 1) / BK;
    const int c_beg = blockIdx.y * p.chunks_per_split;
    const int c_end = min(total_chunks, c_beg + p.chunks_per_split);
    const int nchunks = max(c_end - c_beg, 0);
    int tmem_cols = 32;
    while (tmem_cols < NT) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), TN_PROD_WARPS); mbar_init(empty_bar(s), 1); }
        mbar_init(tfull_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TN_PROD_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (warp < TN_PROD_WARPS) {
        if (nchunks > 0) {
            // ===== producers ==========================================================================================
            // lane = row (feature i / column j) of a 32-row block; the 4 vertices of the warp form the 16-byte k chunk `warp`
            // of that row.  Loads use clamped, always-valid addresses and are zero-masked at store time, so no instruction
            // of the load phase depends on a loaded value and TN_PREFETCH chunks stay in flight.
            float xa[TN_PREFETCH][16];           // [mb 0..3][t 0..3]
            float gb[TN_PREFETCH][4 * NBMAX];    // [nb][t 0..3]
            // Tail rows (the stage inputs are 128 k + 3 wide: a further 128-row UMMA tile would be 98 % padding and would
            // stage G once more): the producers of tile 0 already hold G[v, j] in registers, so they also load the `tail`
            // extra X columns of their vertices (warp-uniform loads) and keep C[Kin + m, j] partial sums on the CUDA cores.
            const int ntail = (blockIdx.x == 0) ? p.tail : 0;
            float xt[TN_PREFETCH][4 * TN_TAIL_MAX];   // [t 0..3][m]
            float tacc[TN_TAIL_MAX][NBMAX];           // [m][nb]: column j = nb * 32 + lane
#pragma unroll
            for (int m = 0; m < TN_TAIL_MAX; ++m)
#pragma unroll
                for (int nb = 0; nb < NBMAX; ++nb) tacc[m][nb] = 0.f;
            auto load_chunk = [&](int c, float (&x)[16], float (&g)[4 * NBMAX], float (&xtail)[4 * TN_TAIL_MAX]) {
                const int v0 = c * BK + warp * 4;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int v = min(v0 + t, p.V - 1);
#pragma unroll
                    for (int m = 0; m < TN_TAIL_MAX; ++m)
                        if (m < ntail) xtail[t * TN_TAIL_MAX + m] = __ldg(p.Xtail + (size_t)v * p.ld_tail + m);
#pragma unroll
                    for (int mb = 0; mb < 4; ++mb)
#ifdef MRB_DIAG_TN_NOLOAD
                        x[mb * 4 + t] = (float)(v + mb);
#else
                        x[mb * 4 + t] = __ldg(p.X + (size_t)v * p.ldx + min(i0 + mb * 32 + lane, p.Kin - 1));
#endif
#pragma unroll
                    for (int nb = 0; nb < NBMAX; ++nb)
#ifdef MRB_DIAG_TN_NOLOAD
                        if (nb * 32 < NT) g[nb * 4 + t] = (float)(v - nb);
#else
                        if (nb * 32 < NT) g[nb * 4 + t] = __ldg(p.G + (size_t)v * p.ldg + min(nb * 32 + lane, NT - 1));
#endif
                }
            };
            // nvalid: how many of the 4 vertices of this 16-byte k chunk exist (the rest is zero padding)
            auto split_store = [&](unsigned char* hi_t, unsigned char* lo_t, int row, const float* v, int nvalid) {
                const int off = (row >> 3) * 1024 + (row & 7) * 128 + (((warp ^ (row & 7)) & 7) << 4);
                const float v0 = nvalid > 0 ? v[0] : 0.f, v1 = nvalid > 1 ? v[1] : 0.f, v2 = nvalid > 2 ? v[2] : 0.f,
                            v3 = nvalid > 3 ? v[3] : 0.f;
                float4 h, l;
                h.x = tf32_rna(v0); h.y = tf32_rna(v1); h.z = tf32_rna(v2); h.w = tf32_rna(v3);
                l.x = tf32_rna(v0 - h.x); l.y = tf32_rna(v1 - h.y); l.z = tf32_rna(v2 - h.z); l.w = tf32_rna(v3 - h.w);
                *reinterpret_cast<float4*>(hi_t + off) = h;
                *reinterpret_cast<float4*>(lo_t + off) = l;
            };
#pragma unroll
            for (int d = 0; d < TN_PREFETCH; ++d)
                if (d < nchunks) load_chunk(c_beg + d, xa[d], gb[d], xt[d]);
            for (int cb = 0; cb < nchunks; cb += TN_PREFETCH) {
#pragma unroll
                for (int d = 0; d < TN_PREFETCH; ++d) {
                    const int c = cb + d;
                    if (c < nchunks) {
                        const int s = c % STAGES, use = c / STAGES;
                        if (use > 0) mbar_wait(empty_bar(s), (use - 1) & 1);
                        if (threadIdx.x == 0) TLY(0, 50);
                        unsigned char* a_hi = smem + s * STAGE_BYTES;
                        unsigned char* a_lo = a_hi + A_BYTES;
                        unsigned char* b_hi = a_hi + 2 * A_BYTES;
                        unsigned char* b_lo = b_hi + b_bytes;
                        const int vleft = p.V - ((c_beg + c) * BK + warp * 4);          // vertices left from this k chunk on
#pragma unroll
                        for (int mb = 0; mb < 4; ++mb)
                            split_store(a_hi, a_lo, mb * 32 + lane, &xa[d][mb * 4], (i0 + mb * 32 + lane < p.Kin) ? vleft : 0);
#pragma unroll
                        for (int nb = 0; nb < NBMAX; ++nb)
                            if (nb * 32 < NT) split_store(b_hi, b_lo, nb * 32 + lane, &gb[d][nb * 4], vleft);
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(full_bar(s));
                        if (threadIdx.x == 0) TLY(0, 51);
                        if (ntail) {
#pragma unroll
                            for (int t = 0; t < 4; ++t)
#pragma unroll
                                for (int m = 0; m < TN_TAIL_MAX; ++m) {
                                    const float xv = (m < ntail && t < vleft) ? xt[d][t * TN_TAIL_MAX + m] : 0.f;
#pragma unroll
                                    for (int nb = 0; nb < NBMAX; ++nb)
                                        if (nb * 32 < NT) tacc[m][nb] = fmaf(xv, gb[d][nb * 4 + t], tacc[m][nb]);
                                }
                        }
                        if (c + TN_PREFETCH < nchunks) load_chunk(c_beg + c + TN_PREFETCH, xa[d], gb[d], xt[d]);
                    }
                }
            }
            // ===== epilogue: TMEM -> registers -> vectorised fp32 reductions into C ========================================
            // warp w reads TMEM lane quadrant w % 4 (rows) and the column half w / 4
            if (threadIdx.x == 0) TLY(0, 46);
            mbar_wait(tfull_bar, 0);
            if (threadIdx.x == 0) TLY(0, 40);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int ew = warp & 3, half = warp >> 2;
            const int i = i0 + ew * 32 + lane;          // TMEM lane = row of the tile
            const int nblk = NT / 32, blk_beg = half * ((nblk + 1) / 2), blk_end = half ? nblk : (nblk + 1) / 2;
            for (int blk = blk_beg; blk < blk_end; ++blk) {
                const int c0 = blk * 32;
                uint32_t v[32];
                tmem_ld32(tmem_d + ((uint32_t)(ew * 32) << 16) + (uint32_t)c0, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#ifdef MRB_DIAG_TN_NOEPI
                if (i < p.Kin && __uint_as_float(v[0]) == 1.2345e-33f) {
#else
                if (i < p.Kin) {
#endif
                    float* dst = (c0 < p.n_split) ? p.C0 + (size_t)i * p.ldc + c0 : p.C1 + (size_t)i * p.ldc + (c0 - p.n_split);
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                                     "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])),
                                     "f"(__uint_as_float(v[j + 3]))
                                     : "memory");
                }
            }
            if (threadIdx.x == 0) TLY(0, 42);
            if (ntail) {
                // all MMAs have retired (tfull), so the stage ring is free: sum the 8 warps' partial tail rows there
                float* red = reinterpret_cast<float*>(smem);                   // [warp][m][256]
                asm volatile("bar.sync 1, %0;" ::"r"(TN_PROD_WARPS * 32));     // every producer warp is past its last stage use
#pragma unroll
                for (int m = 0; m < TN_TAIL_MAX; ++m)
#pragma unroll
                    for (int nb = 0; nb < NBMAX; ++nb)
                        if (m < ntail && nb * 32 < NT) red[(warp * TN_TAIL_MAX + m) * 256 + nb * 32 + lane] = tacc[m][nb];
                asm volatile("bar.sync 1, %0;" ::"r"(TN_PROD_WARPS * 32));
                for (int e = threadIdx.x; e < ntail * NT; e += TN_PROD_WARPS * 32) {
                    const int m = e / NT, j = e - m * NT;
                    float sum = 0.f;
#pragma unroll
                    for (int w = 0; w < TN_PROD_WARPS; ++w) sum += red[(w * TN_TAIL_MAX + m) * 256 + j];
                    float* dst = (j < p.n_split) ? p.T0 + (size_t)m * p.ldc + j : p.T1 + (size_t)m * p.ldc + (j - p.n_split);
                    atomicAdd(dst, sum);
                }
            }
        }
    } else {
        if (lane == 0 && nchunks > 0) {
            const uint32_t idesc = umma_idesc(NT);
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % STAGES, use = c / STAGES;
                mbar_wait(full_bar(s), use & 1);
                TLY(2, 20);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = smem_base + s * STAGE_BYTES, a_lo = a_hi + A_BYTES;
                const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + b_bytes;
#pragma unroll
                for (int kk = 0; kk < BK / 8; ++kk) {
                    const uint64_t dah = umma_desc(a_hi + kk * 32), dal = umma_desc(a_lo + kk * 32);
                    const uint64_t dbh = umma_desc(b_hi + kk * 32), dbl = umma_desc(b_lo + kk * 32);
#ifndef MRB_DIAG_TN_NOMMA
                    umma_tf32(tmem_d, dal, dbh, idesc, (c | kk) != 0);
                    umma_tf32(tmem_d, dah, dbl, idesc, 1);
                    umma_tf32(tmem_d, dah, dbh, idesc, 1);
#endif
                }
                umma_commit(empty_bar(s));
            }
            umma_commit(tfull_bar);
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == TN_PROD_WARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
    }
}

}  // namespace gemmtc
}  // namespace mrb

using namespace mrb;
using namespace mrb::gemmtc;

#ifdef MRB_TC_TIMELINE
extern "C" int mrb_debug_tc_timeline(long long* stamps, int* counts, int reset) {
    if (reset) { int z[4] = {0, 0, 0, 0}; cudaMemcpyToSymbol(g_tl_n, z, sizeof(z)); return 0; }
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(stamps, g_tl, sizeof(long long) * 4 * 2048);
    cudaMemcpyFromSymbol(counts, g_tl_n, sizeof(int) * 4);
    return 0;
}
#endif

extern "C" long long mrb_gemm_tc_image_bytes(int K, int N) {
    if (K <= 0 || N <= 0) return -1;
    return (long long)make_plan(K, N).image_bytes;
}

extern "C" int mrb_gemm_tc_pack(const float* src0, const float* src1, long long stride_k, long long stride_n,
                                int split_axis, int split_at, int K, int N, void* image, void* stream_) {
    MRB_REQUIRE(src0 && image && K > 0 && N > 0, "gemm_tc_pack: bad arguments");
    MRB_REQUIRE(split_axis == 0 || src1, "gemm_tc_pack: second source missing");
    MRB_REQUIRE(((uintptr_t)image & 15) == 0, "gemm_tc_pack: image must be 16-byte aligned");
    const Plan pl = make_plan(K, N);
    PackParams p;
    p.src0 = src0; p.src1 = src1; p.sk = stride_k; p.sn = stride_n; p.split_axis = split_axis; p.split_at = split_at;
    p.K = K; p.N = N; p.NT = pl.NT; p.nchunks = pl.nchunks; p.ntiles = pl.ntiles;
    p.image = (unsigned char*)image;
    const long long total = (long long)pl.ntiles * pl.nchunks * pl.NT * BK;
    k_pack_b<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream_>>>(p);
    return check_launch("gemm_tc_pack");
}

extern "C" int mrb_gemm_tc_pack_graphconv(const float* w0, const float* w1, int K, int D, void* image_fwd, void* image_bwd,
                                          void* stream_) {
    MRB_REQUIRE(w0 && w1 && image_fwd && image_bwd && K > 0 && D > 0, "gemm_tc_pack_graphconv: bad arguments");
    MRB_REQUIRE((((uintptr_t)image_fwd | (uintptr_t)image_bwd) & 15) == 0, "gemm_tc_pack_graphconv: images must be 16-byte aligned");
    // forward operand: B(k, n) = [W0 | W1](k, n), K x 2D;  input-gradient operand: B(k, n) = [W0 | W1](n, k), 2D x K
    const Plan pf = make_plan(K, 2 * D), pb = make_plan(2 * D, K);
    PackParams a, b;
    a.src0 = w0; a.src1 = w1; a.sk = D; a.sn = 1; a.split_axis = 1; a.split_at = D;
    a.K = K; a.N = 2 * D; a.NT = pf.NT; a.nchunks = pf.nchunks; a.ntiles = pf.ntiles; a.image = (unsigned char*)image_fwd;
    b.src0 = w0; b.src1 = w1; b.sk = 1; b.sn = D; b.split_axis = 2; b.split_at = D;
    b.K = 2 * D; b.N = K; b.NT = pb.NT; b.nchunks = pb.nchunks; b.ntiles = pb.ntiles; b.image = (unsigned char*)image_bwd;
    const long long ta = (long long)pf.ntiles * pf.nchunks * pf.NT * BK, tb = (long long)pb.ntiles * pb.nchunks * pb.NT * BK;
    k_pack_b2<<<(unsigned)ceil_div64(ta + tb, 256), 256, 0, (cudaStream_t)stream_>>>(a, b, ta);
    return check_launch("gemm_tc_pack_graphconv");
}

extern "C" int mrb_gemm_tc_pack_graphconv_batch(int n, const void* const* w0, const void* const* w1, const int* K, const int* D,
                                                void* const* image_fwd, void* const* image_bwd, void* stream_) {
    MRB_REQUIRE(n >= 0 && (n == 0 || (w0 && w1 && K && D && image_fwd && image_bwd)), "gemm_tc_pack_graphconv_batch: bad arguments");
    PackBatch b;
    b.n = 0;
    unsigned blocks = 0;
    auto flush = [&]() {
        if (b.n == 0) return;
        b.first_block[b.n] = blocks;
        k_pack_batch<<<blocks, 256, 0, (cudaStream_t)stream_>>>(b);
        b.n = 0;
        blocks = 0;
    };
    auto add = [&](const PackParams& e, const Plan& pl) {
        if (b.n == PACK_BATCH_MAX) flush();
        b.e[b.n] = e;
        b.first_block[b.n] = blocks;
        blocks += (unsigned)ceil_div64((long long)pl.ntiles * pl.nchunks * pl.NT * BK, 256);
        ++b.n;
    };
    for (int i = 0; i < n; ++i) {
        MRB_REQUIRE(w0[i] && w1[i] && image_fwd[i] && K[i] > 0 && D[i] > 0, "gemm_tc_pack_graphconv_batch: bad entry %d", i);
        MRB_REQUIRE((((uintptr_t)image_fwd[i] | (uintptr_t)image_bwd[i]) & 15) == 0,
                    "gemm_tc_pack_graphconv_batch: images must be 16-byte aligned");
        // same operands as mrb_gemm_tc_pack_graphconv: forward [W0 | W1] (K x 2D), input gradient [W0 | W1]^T (2D x K)
        const Plan pf = make_plan(K[i], 2 * D[i]);
        PackParams a;
        a.src0 = (const float*)w0[i]; a.src1 = (const float*)w1[i]; a.sk = D[i]; a.sn = 1; a.split_axis = 1; a.split_at = D[i];
        a.K = K[i]; a.N = 2 * D[i]; a.NT = pf.NT; a.nchunks = pf.nchunks; a.ntiles = pf.ntiles; a.image = (unsigned char*)image_fwd[i];
        add(a, pf);
        if (image_bwd[i]) {
            const Plan pb = make_plan(2 * D[i], K[i]);
            PackParams c;
            c.src0 = (const float*)w0[i]; c.src1 = (const float*)w1[i]; c.sk = 1; c.sn = D[i]; c.split_axis = 2; c.split_at = D[i];
            c.K = 2 * D[i]; c.N = K[i]; c.NT = pb.NT; c.nchunks = pb.nchunks; c.ntiles = pb.ntiles; c.image = (unsigned char*)image_bwd[i];
            add(c, pb);
        }
    }
    flush();
    return check_launch("gemm_tc_pack_graphconv_batch");
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup: the library keeps no link-time dependency on
// libcuda (it must load on machines without a driver, e.g. for the symbol checks of the CPU test suite).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static std::atomic<void*> cached{nullptr};
    void* f = cached.load(std::memory_order_acquire);
    if (!f) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        cached.store(f, std::memory_order_release);
    }
    return (EncodeTiledFn)f;
}

// 2-D fp32 tensor map of C (M rows of ldc floats, N columns used) with 32 x 32 boxes in the SWIZZLE_128B shared-memory layout
static bool make_c_map(CUtensorMap* map, float* C, int M, int N, int ldc) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)ldc * 4};
    const cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, C, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#ifndef MRB_TC_TMA_STORE
#define MRB_TC_TMA_STORE 1
#endif

static int launch_gemm_tc(const float* A, int lda, int M, int K, const void* image, int N, float* C, int ldc, int accumulate,
                          void* stream_, const char* what) {
    MRB_REQUIRE(A && image && C, "%s: null pointer", what);
    MRB_REQUIRE(M >= 0 && K > 0 && N > 0 && lda >= K && ldc >= N, "%s: bad shape M=%d K=%d N=%d lda=%d ldc=%d", what, M, K, N,
                lda, ldc);
    MRB_REQUIRE(((uintptr_t)image & 15) == 0, "%s: image must be 16-byte aligned", what);
    if (M == 0) return MRB_OK;
    static SmemOptIn optin;
    if (int rc = ensure_dynamic_smem(k_gemm_tc, TC_SMEM_LIMIT, optin, what)) return rc;
    const Plan pl = make_plan(K, N);
    // The tensor-core fp32 accumulate truncates, so the error of an accumulator grows with the number of MMAs chained
    // into it.  K is therefore processed in segments of SEG chunks (1024 columns); every segment is a full pass whose
    // epilogue adds into C with ordinary fp32 rounding.
    const int SEG = 32;
    const size_t b_bytes = (size_t)pl.NT * BK * 4;
    alignas(64) CUtensorMap cmap;
    memset(&cmap, 0, sizeof(cmap));
    const bool c_aligned = (ldc % 4 == 0) && (((uintptr_t)C & 15) == 0);
    const bool tma_store = MRB_TC_TMA_STORE && c_aligned && make_c_map(&cmap, C, M, N, ldc);
    for (int c0 = 0; c0 < pl.nchunks; c0 += SEG) {
        const int nch = min(SEG, pl.nchunks - c0);
        Params p;
        p.A = A + (size_t)c0 * BK; p.lda = lda; p.M = M; p.K = min(K - c0 * BK, nch * BK);
        p.image = (const unsigned char*)image + (size_t)c0 * 2 * b_bytes;
        p.img_tile_stride = (long long)pl.nchunks * 2 * (long long)b_bytes;
        p.N = N; p.NT = pl.NT; p.nchunks = nch; p.tmem_cols = pl.tmem_cols;
        // accumulators: one for the lo cross terms + 1 or 3 main ones (long segments, when TMEM allows); power of two
        p.nacc = 2;
        if (nch > 16 && 4 * pl.tmem_cols <= 512) p.nacc = 4;
        // 256-column tiles: main + cross-term accumulators fill the 512 TMEM columns, so the tensor pipe would idle during
        // every epilogue.  For reductions of <= TC_SINGLE_ACC_CHUNKS chunks (<= 192 chained MMAs; the weight-gradient kernel
        // chains 384) one accumulator is accurate enough (measured rel. L2 error vs fp64 in tests/test_gemm_gpu.py) and the
        // accumulator can be double buffered.
        if (2 * p.nacc * pl.tmem_cols > 512 && nch <= TC_SINGLE_ACC_CHUNKS) p.nacc = 1;
        p.acc_bufs = (2 * p.nacc * pl.tmem_cols <= 512) ? 2 : 1;
        p.mtiles = ceil_div(M, BM);
        p.ntiles = pl.ntiles;
        p.accumulate = (c0 > 0) || accumulate;
        p.C = C; p.ldc = ldc; p.tma_store = tma_store;
        p.stages = pl.stages; p.stage_bytes = pl.stage_bytes;
        const int grid = min(p.mtiles * p.ntiles, kNumSMs);
        k_gemm_tc<<<grid, TC_THREADS, pl.stages * pl.stage_bytes + EPI_BYTES + 1024 + 256, (cudaStream_t)stream_>>>(p, cmap);
    }
    return check_launch(what);
}

extern "C" int mrb_gemm_tc(const float* A, int lda, int M, int K, const void* image, int N, float* C, int ldc,
                           void* stream_) {
    return launch_gemm_tc(A, lda, M, K, image, N, C, ldc, 0, stream_, "gemm_tc");
}

extern "C" int mrb_gemm_tc_acc(const float* A, int lda, int M, int K, const void* image, int N, float* C, int ldc,
                               int accumulate, void* stream_) {
    return launch_gemm_tc(A, lda, M, K, image, N, C, ldc, accumulate != 0, stream_, "gemm_tc_acc");
}

static int launch_wgrad(const float* X, int ldx, const float* G, int ldg, int V, int Kin, int N, float* C0, float* C1,
                        int n_split, int ldc, const float* Xtail, int ld_tail, int n_tail, float* T0, float* T1,
                        void* stream_, const char* what) {
    MRB_REQUIRE(X && G && C0, "%s: null pointer", what);
    MRB_REQUIRE(N % 32 == 0 && N >= 32 && N <= NT_MAX, "%s: N must be a multiple of 32 in [32, 256], got %d", what, N);
    MRB_REQUIRE(n_split % 32 == 0 && n_split > 0 && n_split <= N && (n_split == N || C1), "%s: bad column split", what);
    MRB_REQUIRE(ldc % 4 == 0 && ((uintptr_t)C0 & 15) == 0 && (!C1 || ((uintptr_t)C1 & 15) == 0),
                "%s: C rows must be 16-byte aligned", what);
    MRB_REQUIRE(Kin > 0 && V >= 0 && ldx >= Kin && ldg >= N, "%s: bad shape", what);
    MRB_REQUIRE(n_tail >= 0 && n_tail <= TN_TAIL_MAX && (n_tail == 0 || (Xtail && T0 && ld_tail >= n_tail && (n_split == N || T1))),
                "%s: bad tail (0..%d extra columns)", what, TN_TAIL_MAX);
    if (V == 0) return MRB_OK;
    static SmemOptIn optin8, optin4;
    if (int rc = ensure_dynamic_smem(k_gemm_tn<8, 1>, SMEM_BYTES, optin8, what)) return rc;
    if (int rc = ensure_dynamic_smem(k_gemm_tn<4, 2>, SMEM_BYTES, optin4, what)) return rc;
    ParamsTN p;
    p.X = X; p.ldx = ldx; p.G = G; p.ldg = ldg; p.V = V; p.Kin = Kin; p.N = N; p.C0 = C0; p.C1 = C1; p.n_split = n_split;
    p.ldc = ldc; p.tail = n_tail; p.Xtail = Xtail; p.ld_tail = ld_tail; p.T0 = T0; p.T1 = T1;
    const int mtiles = ceil_div(Kin, BM);
    const int total_chunks = ceil_div(V, BK);
    // <= 32 chunks (384 chained, truncating MMAs) per CTA, and a CTA count that fills whole waves of the 148 SMs: with the cap
    // alone, V = 206k gave 202 CTAs = 1.36 waves (a third of the machine idle during the second one)
    const int per_wave = max(1, kNumSMs / mtiles);
    int waves = 1;
    while (ceil_div(total_chunks, per_wave * waves) > 32) ++waves;
    p.chunks_per_split = max(1, ceil_div(total_chunks, per_wave * waves));
    const int splits = ceil_div(total_chunks, p.chunks_per_split);
    // (splitting N = 256 into two 128-column CTAs so that <4, 2> keeps two chunks in flight was measured SLOWER -- 72 vs 59 us
    //  at V = 50k, 247 vs 176 us at 206k: the cost is per chunk iteration, not per byte; profiles/r02/gemm_tn_timeline*.txt)
    if (N <= 128) k_gemm_tn<4, 2><<<dim3(mtiles, splits), TN_THREADS, SMEM_BYTES, (cudaStream_t)stream_>>>(p);
    else k_gemm_tn<8, 1><<<dim3(mtiles, splits), TN_THREADS, SMEM_BYTES, (cudaStream_t)stream_>>>(p);
    return check_launch(what);
}

extern "C" int mrb_gemm_tc_wgrad(const float* X, int ldx, const float* G, int ldg, int V, int Kin, int N, float* C0,
                                 float* C1, int n_split, int ldc, void* stream_) {
    MRB_REQUIRE(X && C0 && Kin > 0, "gemm_tc_wgrad: bad arguments");
    // the 3 position columns of a concatenated stage input (Kin = 128 k + 3): rows [Kin - tail, Kin) go through the tail path
    const int tail = (Kin > BM && Kin % BM <= TN_TAIL_MAX) ? Kin % BM : 0;
    const int Km = Kin - tail;
    return launch_wgrad(X, ldx, G, ldg, V, Km, N, C0, C1, n_split, ldc, X + Km, ldx, tail, C0 + (size_t)Km * ldc,
                        C1 ? C1 + (size_t)Km * ldc : nullptr, stream_, "gemm_tc_wgrad");
}

extern "C" int mrb_gemm_tc_wgrad_split(const float* X, int ldx, const float* G, int ldg, int V, int Kin, int N, float* C0,
                                       float* C1, int n_split, int ldc, const float* Xtail, int ld_tail, int n_tail, float* T0,
                                       float* T1, void* stream_) {
    return launch_wgrad(X, ldx, G, ldg, V, Kin, N, C0, C1, n_split, ldc, Xtail, ld_tail, n_tail, T0, T1, stream_,
                        "gemm_tc_wgrad_split");
}

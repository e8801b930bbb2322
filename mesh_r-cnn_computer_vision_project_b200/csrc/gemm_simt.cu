// Generic fp32 GEMM on the CUDA cores (exact fp32 accumulate): C = op(A) * op(B) + beta * C, row-major.
//
// Used for the small / odd-shaped contractions of the refinement stages (128->3 heads, K = 3 position columns,
// weight gradients) and as the always-available exact-fp32 path; the N = 128/256 GraphConv projections run on the
// tcgen05 tensor cores (gemm_tc.cu).  Replaces the torch.mm / nn.Linear call sites at reference
// meshRCNN/layers.py:54,57,93,155,230,255,335.
//
// 64x64 tile, BK = 16, 256 threads, 4x4 register micro-tile; split-K over gridDim.z with fp32 atomics for the
// tall-skinny weight-gradient shape (K = number of vertices).
#include "common.cuh"
#include "../../include/meshrcnn_b200.h"

namespace mrb {
namespace gemm {

constexpr int BM = 64, BN = 64, BK = 16;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) k_sgemm(int M, int N, int K, const float* __restrict__ A, int lda,
                                               const float* __restrict__ B, int ldb, float beta, float* __restrict__ C,
                                               int ldc, int k_per_split, int atomic_out) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    const int tx = tid & 15, ty = tid >> 4;   // 16 x 16 threads, each 4 x 4 outputs
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        // A tile: BM x BK  (element (m,k): TA ? A[k*lda+m] : A[m*lda+k])
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int idx = tid + r * 256;   // 0..1023
            int m, k;
            if (TA) { m = idx & 63; k = idx >> 6; } else { k = idx & 15; m = idx >> 4; }
            const int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < M && gk < kend) v = TA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
            As[k][m] = v;
        }
        // B tile: BK x BN  (element (k,n): TB ? B[n*ldb+k] : B[k*ldb+n])
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int idx = tid + r * 256;
            int n, k;
            if (TB) { k = idx & 15; n = idx >> 4; } else { n = idx & 63; k = idx >> 6; }
            const int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < N && gk < kend) v = TB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float* c = C + (size_t)gm * ldc + gn;
            if (atomic_out) atomicAdd(c, acc[i][j]);
            else *c = (beta == 0.f) ? acc[i][j] : fmaf(beta, *c, acc[i][j]);
        }
    }
}

__global__ void k_scale_rows(float* __restrict__ C, int M, int N, int ldc, float beta) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)M * N) return;
    float* c = C + (size_t)(t / N) * ldc + (t % N);
    *c = (beta == 0.f) ? 0.f : *c * beta;
}

// ---------------------------------------------------------------------------------------------------------
// skinny shapes of the 3-wide position heads (128 -> 3, 131 -> 3): one of the GEMM dimensions is <= 8
// ---------------------------------------------------------------------------------------------------------
constexpr int SKINNY = 8;

// C[M x N] = A[M x K] * Bop[K x N], N <= NACC, Bop(k, n) = B[n * sbn + k * sbk] (either orientation of the small operand):
// one warp per row (coalesced reads of the long rows), lanes stride over K, two rows per warp in flight, NACC accumulators
// per row, warp reduce
template <int NACC>
__global__ void __launch_bounds__(256) k_skinny_nt(int M, int N, int K, const float* __restrict__ A, int lda,
                                                   const float* __restrict__ B, int sbn, int sbk, float beta,
                                                   float* __restrict__ C, int ldc) {
    const int m0 = (blockIdx.x * (blockDim.x >> 5) + warp_id()) * 2;
    if (m0 >= M) return;
    const int m1 = min(m0 + 1, M - 1);
    float acc0[NACC], acc1[NACC];
#pragma unroll
    for (int n = 0; n < NACC; ++n) acc0[n] = acc1[n] = 0.f;
    for (int k = lane_id(); k < K; k += 32) {
        const float a0 = A[(size_t)m0 * lda + k], a1 = A[(size_t)m1 * lda + k];
#pragma unroll
        for (int n = 0; n < NACC; ++n)
            if (n < N) {
                const float b = __ldg(B + (size_t)n * sbn + (size_t)k * sbk);
                acc0[n] = fmaf(a0, b, acc0[n]);
                acc1[n] = fmaf(a1, b, acc1[n]);
            }
    }
#pragma unroll
    for (int n = 0; n < NACC; ++n)
        if (n < N) { acc0[n] = warp_sum(acc0[n]); acc1[n] = warp_sum(acc1[n]); }
    if (lane_id() < N) {
        float v0 = 0.f, v1 = 0.f;
#pragma unroll
        for (int n = 0; n < NACC; ++n)
            if (n == lane_id()) { v0 = acc0[n]; v1 = acc1[n]; }
        float* c0 = C + (size_t)m0 * ldc + lane_id();
        *c0 = (beta == 0.f) ? v0 : fmaf(beta, *c0, v0);
        if (m0 + 1 < M) {
            float* c1 = C + (size_t)(m0 + 1) * ldc + lane_id();
            *c1 = (beta == 0.f) ? v1 : fmaf(beta, *c1, v1);
        }
    }
}

// C[M x N] = A[M x K] * Bop[K x N], K <= 8, Bop(k, n) = B[k * sbk + n * sbn]: one warp per row, lanes stride over the N
// columns (coalesced stores, the K values of the row are broadcast loads, B comes from L1)
__global__ void __launch_bounds__(256) k_skinny_k(int M, int N, int K, const float* __restrict__ A, int lda,
                                                  const float* __restrict__ B, int sbk, int sbn, float beta,
                                                  float* __restrict__ C, int ldc) {
    const int m = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (m >= M) return;
    float a[SKINNY];
#pragma unroll
    for (int k = 0; k < SKINNY; ++k) a[k] = (k < K) ? A[(size_t)m * lda + k] : 0.f;
    for (int n = lane_id(); n < N; n += 32) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < SKINNY; ++k)
            if (k < K) acc = fmaf(a[k], __ldg(B + (size_t)k * sbk + (size_t)n * sbn), acc);
        float* c = C + (size_t)m * ldc + n;
        *c = (beta == 0.f) ? acc : fmaf(beta, *c, acc);
    }
}

// C[M x N] += A[K x M]^T * B[K x N], M <= MACC (C pre-scaled by beta): a block owns a chunk of the long K dimension; warp w
// takes rows w, w + 8, ... of the chunk, lanes stride over the N <= 256 columns (coalesced), the 8 warps' partial sums are
// combined in shared memory one output row at a time (every index into acc[][] is a compile-time constant, so the
// partial sums stay in registers) and added to C with one atomic per element and block.
// rows in flight per warp / blocks per SM, measured at C[3 x 128] = gpre^T x with 50 k (206 k) rows (scripts/variants.sh gemm_simt.cu "python scripts/time_skinny.py"):
// 2 / 4: 27.3 (66.7) us; 4 / 4: 20.3 (42.0); 4 / 2: 17.9 (41.7); 6 / 2: 17.9 (38.2); 4 / 8: 19.8 (44.4)
#ifndef MRB_SK_TN_ROWS
#define MRB_SK_TN_ROWS 4
#endif
#ifndef MRB_SK_TN_RPB_DIV
#define MRB_SK_TN_RPB_DIV 2
#endif
constexpr int SK_TN_ROWS = MRB_SK_TN_ROWS;
template <int MACC>
__global__ void __launch_bounds__(256) k_skinny_tn(int M, int N, int K, const float* __restrict__ A, int lda,
                                                   const float* __restrict__ B, int ldb, float* __restrict__ C, int scm,
                                                   int scn, int rows_per_block) {
    constexpr int JMAX = 8;                                // N <= 256;  C(m, n) = C[m * scm + n * scn]
    __shared__ float red[8][32 * JMAX];
    const int k0 = blockIdx.x * rows_per_block, k1 = min(K, k0 + rows_per_block);
    float acc[MACC][JMAX];
#pragma unroll
    for (int m = 0; m < MACC; ++m)
#pragma unroll
        for (int j = 0; j < JMAX; ++j) acc[m][j] = 0.f;
    // SK_TN_ROWS rows of the chunk per iteration: their loads are all issued before the first FMA (the loop is latency bound:
    // a warp owns only ~10 rows)
    constexpr int R = SK_TN_ROWS;
    for (int k = k0 + warp_id(); k < k1; k += 8 * R) {
        float a[R][MACC], b[R][JMAX];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int kr = min(k + 8 * r, k1 - 1);
            const bool has = k + 8 * r < k1;
#pragma unroll
            for (int m = 0; m < MACC; ++m) a[r][m] = (m < M && has) ? __ldg(A + (size_t)kr * lda + m) : 0.f;
#pragma unroll
            for (int j = 0; j < JMAX; ++j) {
                const int n = j * 32 + lane_id();
                b[r][j] = (n < N) ? __ldg(B + (size_t)kr * ldb + n) : 0.f;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < JMAX; ++j)
#pragma unroll
                for (int m = 0; m < MACC; ++m) acc[m][j] = fmaf(a[r][m], b[r][j], acc[m][j]);
    }
#pragma unroll
    for (int m = 0; m < MACC; ++m) {
        if (m >= M) break;                                 // uniform
        __syncthreads();
#pragma unroll
        for (int j = 0; j < JMAX; ++j) red[warp_id()][j * 32 + lane_id()] = acc[m][j];
        __syncthreads();
        for (int n = threadIdx.x; n < N; n += blockDim.x) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += red[w][n];
            atomicAdd(C + (size_t)m * scm + (size_t)n * scn, v);
        }
    }
}

// C[M x N] += A[K x M]^T * B[K x N] with M, N <= 4 (e.g. the position block of the head's weight gradient, 3 x 3 over all
// vertices): one thread per row of the long K dimension, 16 register accumulators, warp shuffles + one atomic per element
// and block.  (k_skinny_tn would use 3 of 32 lanes.)
__global__ void __launch_bounds__(256) k_tiny_tn(int M, int N, int K, const float* __restrict__ A, int lda,
                                                 const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc) {
    float acc[4][4];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[m][n] = 0.f;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
        float a[4], b[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) a[m] = (m < M) ? __ldg(A + (size_t)k * lda + m) : 0.f;
#pragma unroll
        for (int n = 0; n < 4; ++n) b[n] = (n < N) ? __ldg(B + (size_t)k * ldb + n) : 0.f;
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int n = 0; n < 4; ++n) acc[m][n] = fmaf(a[m], b[n], acc[m][n]);
    }
    __shared__ float red[8][16];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const float v = warp_sum(acc[m][n]);
            if (lane_id() == 0) red[warp_id()][m * 4 + n] = v;
        }
    __syncthreads();
    if (threadIdx.x < 16) {
        const int m = threadIdx.x >> 2, n = threadIdx.x & 3;
        if (m < M && n < N) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
            atomicAdd(C + (size_t)m * ldc + n, v);
        }
    }
}

}  // namespace gemm
}  // namespace mrb

using namespace mrb;
using namespace mrb::gemm;

extern "C" int mrb_sgemm(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                         float beta, float* C, int ldc, void* stream_) {
    MRB_REQUIRE(A && B && C, "sgemm: null pointer");
    MRB_REQUIRE(M >= 0 && N >= 0 && K >= 0, "sgemm: negative dimension");
    if (M == 0 || N == 0) return MRB_OK;
    cudaStream_t s = (cudaStream_t)stream_;
    if (K > 0 && !transA && N <= SKINNY && K > SKINNY) {            // tall A times a skinny B (either orientation)
        const int sbn = transB ? ldb : 1, sbk = transB ? 1 : ldb;
        if (N <= 4) k_skinny_nt<4><<<ceil_div(M, 16), 256, 0, s>>>(M, N, K, A, lda, B, sbn, sbk, beta, C, ldc);
        else k_skinny_nt<SKINNY><<<ceil_div(M, 16), 256, 0, s>>>(M, N, K, A, lda, B, sbn, sbk, beta, C, ldc);
        return check_launch("sgemm");
    }
    if (K > 0 && !transA && K <= SKINNY) {                          // outer-product-like: K <= 8
        const int sbk = transB ? 1 : ldb, sbn = transB ? ldb : 1;
        k_skinny_k<<<ceil_div(M, 8), 256, 0, s>>>(M, N, K, A, lda, B, sbk, sbn, beta, C, ldc);
        return check_launch("sgemm");
    }
    if (K > 0 && transA && !transB && M <= 4 && N <= 4) {           // reduction over the long K into a tiny C
        k_scale_rows<<<1, 256, 0, s>>>(C, M, N, ldc, beta);
        k_tiny_tn<<<min(ceil_div(K, 256), 2 * kNumSMs), 256, 0, s>>>(M, N, K, A, lda, B, ldb, C, ldc);
        return check_launch("sgemm");
    }
    if (K > 0 && transA && !transB && M <= SKINNY && N <= 256) {    // reduction over the long K into a skinny-row C
        k_scale_rows<<<(unsigned)ceil_div64((long long)M * N, 256), 256, 0, s>>>(C, M, N, ldc, beta);
        const int rpb = max(64, ceil_div(K, MRB_SK_TN_RPB_DIV * kNumSMs));
        if (M <= 4) k_skinny_tn<4><<<ceil_div(K, rpb), 256, 0, s>>>(M, N, K, A, lda, B, ldb, C, ldc, 1, rpb);
        else k_skinny_tn<SKINNY><<<ceil_div(K, rpb), 256, 0, s>>>(M, N, K, A, lda, B, ldb, C, ldc, 1, rpb);
        return check_launch("sgemm");
    }
    if (K > 0 && transA && !transB && N <= SKINNY && M <= 256) {    // same with a skinny-column C: C^T = B^T A
        k_scale_rows<<<(unsigned)ceil_div64((long long)M * N, 256), 256, 0, s>>>(C, M, N, ldc, beta);
        const int rpb = max(64, ceil_div(K, MRB_SK_TN_RPB_DIV * kNumSMs));
        if (N <= 4) k_skinny_tn<4><<<ceil_div(K, rpb), 256, 0, s>>>(N, M, K, B, ldb, A, lda, C, 1, ldc, rpb);
        else k_skinny_tn<SKINNY><<<ceil_div(K, rpb), 256, 0, s>>>(N, M, K, B, ldb, A, lda, C, 1, ldc, rpb);
        return check_launch("sgemm");
    }
    const int gx = ceil_div(N, BN), gy = ceil_div(M, BM);
    // split-K when the output tile grid cannot fill the machine and K is long (weight gradients: K = #vertices)
    int splits = 1;
    if (K >= 4096 && gx * gy < 2 * kNumSMs) {
        splits = min(ceil_div(4 * kNumSMs, gx * gy), ceil_div(K, 512));
        if (splits < 1) splits = 1;
    }
    int kps = ceil_div(ceil_div(K, splits), BK) * BK;
    if (kps == 0) kps = BK;
    splits = K > 0 ? ceil_div(K, kps) : 1;
    const int atomic_out = splits > 1;
    if (atomic_out || K == 0) {
        k_scale_rows<<<(unsigned)ceil_div64((long long)M * N, 256), 256, 0, s>>>(C, M, N, ldc, beta);
        if (K == 0) return check_launch("sgemm");
    }
    dim3 grid(gx, gy, splits);
#define LAUNCH(TA, TB) k_sgemm<TA, TB><<<grid, 256, 0, s>>>(M, N, K, A, lda, B, ldb, beta, C, ldc, kps, atomic_out)
    if (transA) { if (transB) LAUNCH(true, true); else LAUNCH(true, false); }
    else { if (transB) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
    return check_launch("sgemm");
}

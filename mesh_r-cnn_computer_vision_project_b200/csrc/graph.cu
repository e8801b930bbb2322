// Graph kernels: COO -> CSR conversion, CSR neighbour-sum gather (GraphConv aggregation, forward and backward),
// ReLU-mask, segment ids.
//
// Replaces aggregate_neighbours / gen_scatter_params (reference meshRCNN/utils.py:52-97: `matrix[col]` gather +
// `scatter_add_` fp32 atomics) with a deterministic CSR row gather: one warp per vertex row, float4 lanes over
// the feature dimension, neighbours read through L2 (a 53k x 128 fp32 feature matrix is 27 MB, L2 is 126 MB).
// The backward of the aggregation is the same kernel on the transposed CSR (identical arrays for the
// symmetric adjacency Cubify emits).
#include "common.cuh"
#include "../../include/meshrcnn_b200.h"

namespace mrb {
namespace graph {

// ---------------------------------------------------------------------------------------------------------
// generic device-wide exclusive scan for int32 (n up to 2^31): block sums -> scan of sums -> add back
// ---------------------------------------------------------------------------------------------------------
constexpr int SCAN_BLOCK = 1024;

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_blocks(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                            int32_t* __restrict__ sums, int n) {
    __shared__ int scratch[33];
    const int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const int v = (i < n) ? in[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, scratch, &total);
    if (i < n) out[i] = ex;
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_sums(int32_t* __restrict__ sums, int nb) {
    __shared__ int scratch[33];
    int carry = 0;
    for (int base = 0; base < nb; base += SCAN_BLOCK) {
        const int i = base + threadIdx.x;
        const int v = (i < nb) ? sums[i] : 0;
        int total;
        const int ex = block_exclusive_scan(v, scratch, &total);
        if (i < nb) sums[i] = carry + ex;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[nb] = carry;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_add(int32_t* __restrict__ out, const int32_t* __restrict__ sums,
                                                         int n, int nb) {
    const int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    if (i < n) out[i] += sums[blockIdx.x];
    if (i == 0) out[n] = sums[nb];   // total at out[n]
}

// out[0..n] = exclusive scan of in[0..n-1] (out[n] = total). sums: >= ceil(n/1024)+1 ints.
static void exclusive_scan(const int32_t* in, int32_t* out, int32_t* sums, int n, cudaStream_t s) {
    const int nb = ceil_div(n, SCAN_BLOCK);
    k_scan_blocks<<<nb, SCAN_BLOCK, 0, s>>>(in, out, sums, n);
    k_scan_sums<<<1, SCAN_BLOCK, 0, s>>>(sums, nb);
    k_scan_add<<<nb, SCAN_BLOCK, 0, s>>>(out, sums, n, nb);
}

// ---------------------------------------------------------------------------------------------------------
// COO -> CSR
// ---------------------------------------------------------------------------------------------------------
__global__ void k_coo_hist(const long long* __restrict__ rows, int E, int n, int32_t* __restrict__ counts,
                           int32_t* __restrict__ flags) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const long long r = rows[e];
    if (r < 0 || r >= n) { atomicOr(flags + 1, 1); return; }   // out-of-range index: reported via flags[1]
    atomicAdd(counts + r, 1);
    if (e > 0 && rows[e - 1] > r) atomicOr(flags, 1);           // not row-sorted
}

__global__ void k_coo_fill(const long long* __restrict__ rows, const long long* __restrict__ cols, int E, int n,
                           const int32_t* __restrict__ rowptr, int32_t* __restrict__ cursor,
                           int32_t* __restrict__ flags, int32_t* __restrict__ col) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const long long r = rows[e];
    if (r < 0 || r >= n) return;
    long long c = cols[e];
    if (c < 0 || c >= n) { atomicOr(flags + 1, 1); c = r; }   // out-of-range column: flagged like a bad row; keep memory safe
    if (flags[0] == 0) {
        col[e] = (int32_t)c;      // already row-sorted: CSR order == COO order (deterministic)
    } else {
        const int p = atomicAdd(cursor + r, 1);
        col[rowptr[r] + p] = (int32_t)c;
    }
}

// ---------------------------------------------------------------------------------------------------------
// CSR neighbour gather:  out[i,:] = act( self[i,:] + sum_{j in N(i)} nbr[col[j],:] )
// ---------------------------------------------------------------------------------------------------------
template <bool RELU>
__global__ void __launch_bounds__(256) k_gather_vec4(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                     int n, const float* __restrict__ self, int ld_self,
                                                     const float* __restrict__ nbr, int ld_nbr, int D,
                                                     float* __restrict__ out, int ld_out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (row >= n) return;
    const int beg = rowptr[row], end = rowptr[row + 1];
    for (int d = lane_id() * 4; d < D; d += 128) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (self) acc = *reinterpret_cast<const float4*>(self + (size_t)row * ld_self + d);
        int j = beg;
        for (; j + 1 < end; j += 2) {   // two independent loads in flight
            const int c0 = col[j], c1 = col[j + 1];
            const float4 a = __ldg(reinterpret_cast<const float4*>(nbr + (size_t)c0 * ld_nbr + d));
            const float4 b = __ldg(reinterpret_cast<const float4*>(nbr + (size_t)c1 * ld_nbr + d));
            acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
            acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
        }
        if (j < end) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(nbr + (size_t)col[j] * ld_nbr + d));
            acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
        }
        if (RELU) {
            acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
        }
        *reinterpret_cast<float4*>(out + (size_t)row * ld_out + d) = acc;
    }
}

template <bool RELU>
__global__ void __launch_bounds__(256) k_gather_scalar(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                       int n, const float* __restrict__ self, int ld_self,
                                                       const float* __restrict__ nbr, int ld_nbr, int D,
                                                       float* __restrict__ out, int ld_out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * D) return;
    const int row = (int)(t / D), d = (int)(t % D);
    float acc = self ? self[(size_t)row * ld_self + d] : 0.f;
    for (int j = rowptr[row]; j < rowptr[row + 1]; ++j) acc += __ldg(nbr + (size_t)col[j] * ld_nbr + d);
    if (RELU) acc = fmaxf(acc, 0.f);
    out[(size_t)row * ld_out + d] = acc;
}

// gz = gout * (act_out > 0), written with a leading dimension (first half of the [gY0 | gY1] buffer)
__global__ void k_relu_mask(const float* __restrict__ gout, int ld_g, const float* __restrict__ act, int ld_a,
                            int n, int D, float* __restrict__ gz, int ld_z) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * D) return;
    const int row = (int)(t / D), d = (int)(t % D);
    const float a = act[(size_t)row * ld_a + d];
    gz[(size_t)row * ld_z + d] = a > 0.f ? gout[(size_t)row * ld_g + d] : 0.f;
}

// Backward of relu(self + sum_nbr): gz = gout * (act > 0) is formed on the fly, written to gy[:, 0:D] and gathered over the
// transposed adjacency into gy[:, D:2D] -- one pass instead of a mask kernel plus a gather kernel.
template <bool VEC>
__global__ void __launch_bounds__(256) k_gather_bwd(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int n,
                                                    const float* __restrict__ gout, int ld_g, const float* __restrict__ act,
                                                    int ld_a, int D, float* __restrict__ gy) {
    const int row = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (row >= n) return;
    const int beg = rowptr[row], end = rowptr[row + 1];
    auto masked = [&](int r, int d) {
        float4 g, a;
        if (VEC) {
            g = __ldg(reinterpret_cast<const float4*>(gout + (size_t)r * ld_g + d));
        } else {
            const float* gp = gout + (size_t)r * ld_g + d;
            g = make_float4(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2), __ldg(gp + 3));
        }
        a = __ldg(reinterpret_cast<const float4*>(act + (size_t)r * ld_a + d));
        return make_float4(a.x > 0.f ? g.x : 0.f, a.y > 0.f ? g.y : 0.f, a.z > 0.f ? g.z : 0.f, a.w > 0.f ? g.w : 0.f);
    };
    for (int d = lane_id() * 4; d < D; d += 128) {
        *reinterpret_cast<float4*>(gy + (size_t)row * 2 * D + d) = masked(row, d);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int j = beg;
        for (; j + 1 < end; j += 2) {
            const float4 x = masked(col[j], d), y = masked(col[j + 1], d);
            acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
            acc.x += y.x; acc.y += y.y; acc.z += y.z; acc.w += y.w;
        }
        if (j < end) {
            const float4 x = masked(col[j], d);
            acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
        }
        *reinterpret_cast<float4*>(gy + (size_t)row * 2 * D + D + d) = acc;
    }
}

__global__ void k_segment_ids(const int32_t* __restrict__ offsets, int nseg, int n, int32_t* __restrict__ ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = nseg;            // largest s with offsets[s] <= i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (offsets[mid] <= i) lo = mid; else hi = mid;
    }
    ids[i] = lo;
}

// out[r, :] = [s0[r, :w0] | s1[r, :w1] | s2[r, :w2] | 0 ...] with a row pitch of ld_out >= w0 + w1 + w2 floats
// (torch.cat of the stage inputs, reference layers.py:160-165,241-252,321-334, written with 16-byte aligned rows)
__global__ void __launch_bounds__(256) k_concat_cols(const float* __restrict__ s0, int w0, int ld0, const float* __restrict__ s1,
                                                     int w1, int ld1, const float* __restrict__ s2, int w2, int ld2, int n,
                                                     float* __restrict__ out, int ld_out) {
    // flat element index over a block's row range: fully coalesced stores, 4 independent loads in flight per thread
    constexpr int U = 4;
    const int rows_per_block = (n + gridDim.x - 1) / gridDim.x;
    const int r_beg = blockIdx.x * rows_per_block, r_end = min(n, r_beg + rows_per_block);
    const unsigned total = (unsigned)max(r_end - r_beg, 0) * (unsigned)ld_out;      // < 2^31: blocks cover few rows
    const float* b0 = s0 + (size_t)r_beg * ld0;
    const float* b1 = s1 + (size_t)r_beg * ld1;
    const float* b2 = s2 + (size_t)r_beg * ld2;
    float* o = out + (size_t)r_beg * ld_out;
    for (unsigned e0 = threadIdx.x; e0 < total; e0 += U * blockDim.x) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned e = e0 + u * blockDim.x;
            v[u] = 0.f;
            if (e < total) {
                const unsigned r = e / (unsigned)ld_out, c = e - r * (unsigned)ld_out;
                if ((int)c < w0) v[u] = b0[(size_t)r * ld0 + c];
                else if ((int)c < w0 + w1) v[u] = b1[(size_t)r * ld1 + (c - w0)];
                else if ((int)c < w0 + w1 + w2) v[u] = b2[(size_t)r * ld2 + (c - w0 - w1)];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned e = e0 + u * blockDim.x;
            if (e < total) o[e] = v[u];
        }
    }
}

}  // namespace graph
}  // namespace mrb

using namespace mrb;
using namespace mrb::graph;

extern "C" int mrb_coo_to_csr(const long long* adj, long long E, int n, int transpose, int32_t* rowptr, int32_t* col,
                              int32_t* workspace, void* stream_) {
    MRB_REQUIRE(rowptr && col && workspace && n >= 0 && E >= 0, "coo_to_csr: bad arguments");
    MRB_REQUIRE(E < (1LL << 31), "coo_to_csr: too many edges for int32 CSR");
    cudaStream_t s = (cudaStream_t)stream_;
    const long long* rows = transpose ? adj + E : adj;
    const long long* cols = transpose ? adj : adj + E;
    int32_t* counts = workspace;               // [n+1] histogram, later reused as per-row cursor
    int32_t* flags = workspace + (n + 1);      // [2]
    int32_t* sums = workspace + (n + 1) + 2;   // [ceil(n/1024)+1]  (fits: workspace holds 2*(n+1)+2 ints)
    cudaMemsetAsync(workspace, 0, sizeof(int32_t) * ((size_t)n + 3), s);
    if (E > 0) k_coo_hist<<<ceil_div(E, 256), 256, 0, s>>>(rows, (int)E, n, counts, flags);
    if (n > 0) exclusive_scan(counts, rowptr, sums, n, s);
    else cudaMemsetAsync(rowptr, 0, sizeof(int32_t), s);
    cudaMemsetAsync(counts, 0, sizeof(int32_t) * ((size_t)n + 1), s);
    if (E > 0) k_coo_fill<<<ceil_div(E, 256), 256, 0, s>>>(rows, cols, (int)E, n, rowptr, counts, flags, col);
    return check_launch("coo_to_csr");
}

extern "C" int mrb_csr_gather_fwd(const int32_t* rowptr, const int32_t* col, int n, const float* self, int ld_self,
                                  const float* nbr, int ld_nbr, int D, int relu, float* out, int ld_out, void* stream_) {
    MRB_REQUIRE(rowptr && col && nbr && out && n >= 0 && D > 0, "csr_gather_fwd: bad arguments");
    if (n == 0) return MRB_OK;
    cudaStream_t s = (cudaStream_t)stream_;
    const bool vec = (D % 4 == 0) && (ld_nbr % 4 == 0) && (ld_out % 4 == 0) && (!self || ld_self % 4 == 0) &&
                     ((uintptr_t)nbr % 16 == 0) && ((uintptr_t)out % 16 == 0) && (!self || (uintptr_t)self % 16 == 0);
    if (vec) {
        const int rows_per_block = 8;
        dim3 grid(ceil_div(n, rows_per_block));
        if (relu) k_gather_vec4<true><<<grid, 256, 0, s>>>(rowptr, col, n, self, ld_self, nbr, ld_nbr, D, out, ld_out);
        else k_gather_vec4<false><<<grid, 256, 0, s>>>(rowptr, col, n, self, ld_self, nbr, ld_nbr, D, out, ld_out);
    } else {
        dim3 grid((unsigned)ceil_div64((long long)n * D, 256));
        if (relu) k_gather_scalar<true><<<grid, 256, 0, s>>>(rowptr, col, n, self, ld_self, nbr, ld_nbr, D, out, ld_out);
        else k_gather_scalar<false><<<grid, 256, 0, s>>>(rowptr, col, n, self, ld_self, nbr, ld_nbr, D, out, ld_out);
    }
    return check_launch("csr_gather_fwd");
}

extern "C" int mrb_relu_mask(const float* gout, int ld_g, const float* act, int ld_a, int n, int D, float* gz, int ld_z,
                             void* stream_) {
    MRB_REQUIRE(gout && act && gz, "relu_mask: null pointer");
    if (n == 0 || D == 0) return MRB_OK;
    k_relu_mask<<<(unsigned)ceil_div64((long long)n * D, 256), 256, 0, (cudaStream_t)stream_>>>(gout, ld_g, act, ld_a, n,
                                                                                               D, gz, ld_z);
    return check_launch("relu_mask");
}

extern "C" int mrb_segment_ids(const int32_t* offsets, int nseg, int n, int32_t* ids, void* stream_) {
    MRB_REQUIRE(offsets && ids && nseg > 0, "segment_ids: bad arguments");
    if (n == 0) return MRB_OK;
    k_segment_ids<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream_>>>(offsets, nseg, n, ids);
    return check_launch("segment_ids");
}

extern "C" int mrb_graphconv_bwd_gather(const int32_t* rowptr_t, const int32_t* col_t, int n, const float* gout, int ld_g,
                                        const float* act, int ld_a, int D, float* gy, void* stream_) {
    MRB_REQUIRE(rowptr_t && col_t && gout && act && gy, "graphconv_bwd_gather: null pointer");
    MRB_REQUIRE(D % 4 == 0 && ld_a % 4 == 0 && ((uintptr_t)act % 16 == 0) && ((uintptr_t)gy % 16 == 0),
                "graphconv_bwd_gather: D and the activation rows must be 16-byte aligned");
    if (n == 0) return MRB_OK;
    cudaStream_t s = (cudaStream_t)stream_;
    const bool vec = (ld_g % 4 == 0) && ((uintptr_t)gout % 16 == 0);
    if (vec) k_gather_bwd<true><<<ceil_div(n, 8), 256, 0, s>>>(rowptr_t, col_t, n, gout, ld_g, act, ld_a, D, gy);
    else k_gather_bwd<false><<<ceil_div(n, 8), 256, 0, s>>>(rowptr_t, col_t, n, gout, ld_g, act, ld_a, D, gy);
    return check_launch("graphconv_bwd_gather");
}

extern "C" int mrb_concat_cols(const float* s0, int w0, int ld0, const float* s1, int w1, int ld1, const float* s2, int w2,
                               int ld2, int n, float* out, int ld_out, void* stream_) {
    MRB_REQUIRE(out && s0 && w0 > 0 && (w1 == 0 || s1) && (w2 == 0 || s2), "concat_cols: null pointer");
    MRB_REQUIRE(w1 >= 0 && w2 >= 0 && ld0 >= w0 && ld1 >= w1 && ld2 >= w2 && ld_out >= w0 + w1 + w2, "concat_cols: bad widths");
    if (n <= 0) return MRB_OK;
    const int blocks = (int)min((long long)16 * kNumSMs, ceil_div64(n, 4));      // >= 4 rows per block
    k_concat_cols<<<blocks, 256, 0, (cudaStream_t)stream_>>>(s0, w0, ld0, s1, w1, ld1, s2, w2, ld2, n, out, ld_out);
    return check_launch("concat_cols");
}

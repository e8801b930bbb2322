// "Split-input" GraphConv: the stage inputs of the reference are column concatenations [features | position | aligned]
// (meshRCNN/layers.py:160-165,241-252,321-334) that feed  relu(x W0 + A (x W1))  (layers.py:47-68).  A product with a
// concatenation is a sum of products, so nothing is concatenated here:
//
//     z_i = y0_i + p_i Wp0 + T0[tex_i]  +  sum_{j in N(i)} ( y1_j + T1[tex_j] )  +  ( sum_{j in N(i)} p_j ) Wp1
//
//   y = x_main [W0_x | W1_x]         the dense 128-wide part on the tensor cores (gemm_tc.cu), K a multiple of 4 -> always
//                                    the 16-byte vector producers, no padded K chunk for the 3 position columns
//   p_i Wp                           the 3 position columns: 2 x 12 FMAs per lane in THIS kernel's epilogue (the neighbour
//                                    positions are summed first: 12 extra bytes per neighbour)
//   T = texel_rows [W0_a | W1_a]     VertexAlign (x) the aligned-feature rows of W: the n_img * H * W texels are projected
//                                    once (tiny GEMM) and tex_i is the texel a vertex projects to (-1 = masked): the
//                                    SV x 256 VertexAlign output is never formed (same algebra as align_proj.cu)
//
// The ReLU mask is written as bits (D / 32 words per vertex) and is all the backward needs of this layer's output:
// the fused backward gather reads 16 bytes of upstream gradient + half a byte of mask per (neighbour, lane) instead of
// 32 bytes.  Optional residual input (ResGraphConv skip, layers.py:96-100) is added after the ReLU.
//
// Also here: the fused 3-wide position head  new_pos = pos + tanh([pos | x] W^T)  (layers.py:255-259,335-339) and the
// vertex -> texel-row table.
#include "common.cuh"
#include "valign.cuh"
#include "../../include/meshrcnn_b200.h"

namespace mrb {
namespace gc2 {

__device__ __forceinline__ void add4(float4& a, const float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

struct FwdParams {
    const int32_t* rowptr; const int32_t* col; int n; int D;
    const float* y; int ld_y;                 // [n x 2D]: self half | neighbour half (HAS_Y)
    const float* pos; const float* wp0; const float* wp1;   // pos [n x 3]; wp0 / wp1: 3 x D row-major blocks of W0 / W1 (HAS_POS)
    const int32_t* texrow; const float* T;    // texrow [n]; T [R x 2D] (HAS_TEX)
    int relu; uint32_t* mask;                 // mask: n x ceil(D/32) words (optional)
    const float* residual; int ld_res;        // optional, added after the ReLU
    float* out; int ld_out;
};

#ifndef GC2_FWD_MINB
#define GC2_FWD_MINB 8      // CTAs per SM the forward gather is compiled for (8 x 256 threads = full occupancy at 32 registers)
#endif
#ifndef GC2_FWD_MINB_POS
#define GC2_FWD_MINB_POS 6  // the variants with the position term need ~40 registers
#endif
#ifndef GC2_WAVES
#define GC2_WAVES 4         // CTAs per resident slot: each walks n / (148 * MINB * WAVES) contiguous rows
#endif
#ifndef GC2_BWD_MINB
#define GC2_BWD_MINB 6
#endif

// One warp per vertex row.  The row's neighbour ids are fetched with ONE coalesced load (lane = neighbour; rows of > 32
// neighbours loop) and broadcast with shuffles, so the dependent chain is rowptr -> ids -> rows instead of one extra
// global load per neighbour, and four neighbour rows are in flight per lane.
template <bool HAS_Y, bool HAS_POS, bool HAS_TEX>
__global__ void __launch_bounds__(256, (HAS_POS && HAS_Y) ? GC2_FWD_MINB_POS : GC2_FWD_MINB) k_gather_fwd(const __grid_constant__ FwdParams p) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = lane_id();
    const int D = p.D;
    const float* __restrict__ y = p.y;
    const float* __restrict__ T = p.T;
    // A CTA walks a CONTIGUOUS range of rows, its 8 warps side by side: vertices are numbered in lattice order (b, z, y, x), so
    // consecutive rows share most of their neighbours and the rows a warp gathers were just brought into L1 by its siblings.
    const int rows_per_cta = ((p.n + (int)gridDim.x - 1) / (int)gridDim.x + 7) & ~7;
    const int row_end = min(p.n, ((int)blockIdx.x + 1) * rows_per_cta);
    for (int row = (int)blockIdx.x * rows_per_cta + warp_id(); row < row_end; row += 8) {
    const int beg = __ldg(p.rowptr + row), end = __ldg(p.rowptr + row + 1);
    float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f;     // this lane's share of the sum of the neighbour positions
    float q0 = 0.f, q1 = 0.f, q2 = 0.f;        // own position
    if (HAS_POS) {
        q0 = __ldg(p.pos + 3 * (size_t)row); q1 = __ldg(p.pos + 3 * (size_t)row + 1); q2 = __ldg(p.pos + 3 * (size_t)row + 2);
    }
    const int my_tex = HAS_TEX ? __ldg(p.texrow + row) : -1;
    for (int d0 = 0; d0 < D; d0 += 128) {      // warp-uniform trip count (the mask bits are combined with shuffles)
        const bool active = d0 + lane * 4 < D;
        const int d = active ? d0 + lane * 4 : 0;   // inactive lanes (D % 128 != 0) recompute column block 0 and store nothing
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (HAS_Y) acc = __ldg(reinterpret_cast<const float4*>(y + (size_t)row * p.ld_y + d));
        if (HAS_TEX && my_tex >= 0) add4(acc, __ldg(reinterpret_cast<const float4*>(T + (size_t)my_tex * 2 * D + d)));
        for (int base = beg; base < end; base += 32) {
            const int cnt = min(32, end - base);
            const int nb = lane < cnt ? __ldg(p.col + base + lane) : 0;              // lane j holds neighbour j
            const int nb_tex = (HAS_TEX && lane < cnt) ? __ldg(p.texrow + nb) : -1;
            if (HAS_POS && d0 == 0 && lane < cnt) {      // issued before the row loads below, consumed after them
                ps0 += __ldg(p.pos + 3 * (size_t)nb); ps1 += __ldg(p.pos + 3 * (size_t)nb + 1); ps2 += __ldg(p.pos + 3 * (size_t)nb + 2);
            }
            int j = 0;
            if (HAS_Y) {
                for (; j + 4 <= cnt; j += 4) {
                    const int c0 = __shfl_sync(FULL, nb, j), c1 = __shfl_sync(FULL, nb, j + 1), c2 = __shfl_sync(FULL, nb, j + 2),
                              c3 = __shfl_sync(FULL, nb, j + 3);
                    const float4 a = __ldg(reinterpret_cast<const float4*>(y + (size_t)c0 * p.ld_y + D + d));
                    const float4 b = __ldg(reinterpret_cast<const float4*>(y + (size_t)c1 * p.ld_y + D + d));
                    const float4 c = __ldg(reinterpret_cast<const float4*>(y + (size_t)c2 * p.ld_y + D + d));
                    const float4 e = __ldg(reinterpret_cast<const float4*>(y + (size_t)c3 * p.ld_y + D + d));
                    add4(acc, a); add4(acc, b); add4(acc, c); add4(acc, e);
                }
                for (; j < cnt; ++j) {
                    const int c0 = __shfl_sync(FULL, nb, j);
                    add4(acc, __ldg(reinterpret_cast<const float4*>(y + (size_t)c0 * p.ld_y + D + d)));
                }
            }
            if (HAS_TEX) {
                unsigned have = __ballot_sync(FULL, nb_tex >= 0);                    // neighbours that project into the map
                while (have) {
                    const int src = __ffs(have) - 1;
                    have &= have - 1;
                    const int t = __shfl_sync(FULL, nb_tex, src);
                    add4(acc, __ldg(reinterpret_cast<const float4*>(T + (size_t)t * 2 * D + D + d)));
                }
            }
        }
        if (HAS_POS) {
            if (d0 == 0) { ps0 = warp_sum(ps0); ps1 = warp_sum(ps1); ps2 = warp_sum(ps2); }
            const float qq[3] = {q0, q1, q2}, pp[3] = {ps0, ps1, ps2};
#pragma unroll
            for (int r = 0; r < 3; ++r) {       // one weight row at a time: few live registers
                const float4 a = __ldg(reinterpret_cast<const float4*>(p.wp0 + r * D + d));
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.wp1 + r * D + d));
                acc.x = fmaf(qq[r], a.x, fmaf(pp[r], b.x, acc.x));
                acc.y = fmaf(qq[r], a.y, fmaf(pp[r], b.y, acc.y));
                acc.z = fmaf(qq[r], a.z, fmaf(pp[r], b.z, acc.z));
                acc.w = fmaf(qq[r], a.w, fmaf(pp[r], b.w, acc.w));
            }
        }
        if (p.mask) {
            // bit (d + e) of the row's mask = [z > 0]; lane l owns the nibble at bits 4 * (l % 8) of word d / 32
            unsigned nib = (acc.x > 0.f ? 1u : 0u) | (acc.y > 0.f ? 2u : 0u) | (acc.z > 0.f ? 4u : 0u) | (acc.w > 0.f ? 8u : 0u);
            nib = active ? nib << (4 * (lane & 7)) : 0u;
            nib |= __shfl_xor_sync(FULL, nib, 1);
            nib |= __shfl_xor_sync(FULL, nib, 2);
            nib |= __shfl_xor_sync(FULL, nib, 4);
            if (active && (lane & 7) == 0) p.mask[(size_t)row * ((D + 31) >> 5) + (d >> 5)] = nib;
        }
        if (p.relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
        if (p.residual) add4(acc, __ldg(reinterpret_cast<const float4*>(p.residual + (size_t)row * p.ld_res + d)));
        if (active) *reinterpret_cast<float4*>(p.out + (size_t)row * p.ld_out + d) = acc;
    }
    }
}

struct BwdParams {
    const int32_t* rowptr_t; const int32_t* col_t; int n; int D;
    const float* gout; int ld_g;              // upstream gradient of the layer output (may be a strided view)
    const uint32_t* mask;                     // n x ceil(D/32) words; NULL = no ReLU (all ones)
    float* gy;                                // [n x 2D] = [gz | A^T gz]
    const float* wp0; const float* wp1; float* gpos;   // optional: gpos[i, :] = gz_i Wp0^T + (A^T gz)_i Wp1^T   (n x 3, overwritten)
    const int32_t* texrow; float* gT;         // optional: gT[tex_i, :] += gy_i (zero-filled by the host wrapper)
};

template <bool VEC, bool HAS_POS>
__global__ void __launch_bounds__(256, GC2_BWD_MINB) k_gather_bwd(const __grid_constant__ BwdParams p) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = lane_id();
    const int D = p.D, W = (D + 31) >> 5;
    const float* __restrict__ gout = p.gout;
    const uint32_t* __restrict__ mask = p.mask;
    const int rows_per_cta = ((p.n + (int)gridDim.x - 1) / (int)gridDim.x + 7) & ~7;     // contiguous row range per CTA (see forward)
    const int row_end = min(p.n, ((int)blockIdx.x + 1) * rows_per_cta);
    for (int row = (int)blockIdx.x * rows_per_cta + warp_id(); row < row_end; row += 8) {
    const int beg = __ldg(p.rowptr_t + row), end = __ldg(p.rowptr_t + row + 1);
    float gp0 = 0.f, gp1 = 0.f, gp2 = 0.f;
    auto masked = [&](int r, int d) {
        float4 g;
        const float* gsrc = gout + (size_t)r * p.ld_g + d;
        if (VEC) g = __ldg(reinterpret_cast<const float4*>(gsrc));
        else g = make_float4(__ldg(gsrc), __ldg(gsrc + 1), __ldg(gsrc + 2), __ldg(gsrc + 3));
        if (mask) {
            const unsigned m = __ldg(mask + (size_t)r * W + (d >> 5)) >> (4 * (lane & 7));
            g.x = (m & 1u) ? g.x : 0.f; g.y = (m & 2u) ? g.y : 0.f; g.z = (m & 4u) ? g.z : 0.f; g.w = (m & 8u) ? g.w : 0.f;
        }
        return g;
    };
    const int my_tex = p.texrow ? __ldg(p.texrow + row) : -1;
    for (int d0 = 0; d0 < D; d0 += 128) {      // warp-uniform trip count: the neighbour ids travel through shuffles
        const bool active = d0 + lane * 4 < D;
        const int d = active ? d0 + lane * 4 : 0;   // inactive lanes (D % 128 != 0) recompute column block 0 and store nothing
        const float4 self = masked(row, d);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int base = beg; base < end; base += 32) {
            const int cnt = min(32, end - base);
            const int nb = lane < cnt ? __ldg(p.col_t + base + lane) : 0;             // lane j holds neighbour j
            int j = 0;
            for (; j + 2 <= cnt; j += 2) {
                const int c0 = __shfl_sync(FULL, nb, j), c1 = __shfl_sync(FULL, nb, j + 1);
                const float4 x = masked(c0, d), yv = masked(c1, d);
                add4(acc, x); add4(acc, yv);
            }
            if (j < cnt) add4(acc, masked(__shfl_sync(FULL, nb, j), d));
        }
        if (active) {
            *reinterpret_cast<float4*>(p.gy + (size_t)row * 2 * D + d) = self;
            *reinterpret_cast<float4*>(p.gy + (size_t)row * 2 * D + D + d) = acc;
        }
        if (HAS_POS && active) {
            float g3[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(p.wp0 + r * D + d));
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.wp1 + r * D + d));
                g3[r] = self.x * a.x + self.y * a.y + self.z * a.z + self.w * a.w + acc.x * b.x + acc.y * b.y + acc.z * b.z + acc.w * b.w;
            }
            gp0 += g3[0]; gp1 += g3[1]; gp2 += g3[2];
        }
        if (my_tex >= 0 && active) {
            float* t = p.gT + (size_t)my_tex * 2 * D + d;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(t), "f"(self.x), "f"(self.y), "f"(self.z), "f"(self.w) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(t + D), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
        }
    }
    if (HAS_POS) {
        gp0 = warp_sum(gp0); gp1 = warp_sum(gp1); gp2 = warp_sum(gp2);
        if (lane == 0) { p.gpos[3 * (size_t)row] = gp0; p.gpos[3 * (size_t)row + 1] = gp1; p.gpos[3 * (size_t)row + 2] = gp2; }
    }
    }
}

__global__ void __launch_bounds__(256) k_texrows(const float* __restrict__ pos, const int32_t* __restrict__ vert_mesh,
                                                 const int32_t* __restrict__ mesh_info, int SV, int size, int32_t* __restrict__ texrow) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= SV) return;
    const valign::Texel t = valign::project(pos, vert_mesh, mesh_info, v, size, size);
    texrow[v] = t.valid ? t.img * size * size + t.xy : -1;
}

// ---------------------------------------------------------------------------------------------------------
// 3-wide position head: pre = x Wx^T + pos Wp^T ; delta = tanh(pre) ; new_pos = pos + delta
// W is nn.Linear's out x in matrix (3 x Kin, row pitch ld_w); Wx = W[:, x_col .. x_col + Kx), Wp = W[:, p_col .. p_col + 3).
// One warp per two rows, lanes stride over Kx (coalesced reads of the long rows), warp reduction.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_head_fwd(const float* __restrict__ x, int ld_x, int Kx, const float* __restrict__ pos,
                                                  const float* __restrict__ W, int ld_w, int x_col, int p_col, int has_p, int n,
                                                  float* __restrict__ new_pos, float* __restrict__ delta) {
    const int m0 = (blockIdx.x * (blockDim.x >> 5) + warp_id()) * 2;
    if (m0 >= n) return;
    const int m1 = min(m0 + 1, n - 1);
    float a0[3] = {0.f, 0.f, 0.f}, a1[3] = {0.f, 0.f, 0.f};
    for (int k = lane_id(); k < Kx; k += 32) {
        const float v0 = x[(size_t)m0 * ld_x + k], v1 = x[(size_t)m1 * ld_x + k];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float w = __ldg(W + (size_t)c * ld_w + x_col + k);
            a0[c] = fmaf(v0, w, a0[c]);
            a1[c] = fmaf(v1, w, a1[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { a0[c] = warp_sum(a0[c]); a1[c] = warp_sum(a1[c]); }
    if (lane_id() < 3) {
        const int c = lane_id();
        float s0 = c == 0 ? a0[0] : (c == 1 ? a0[1] : a0[2]);
        float s1 = c == 0 ? a1[0] : (c == 1 ? a1[1] : a1[2]);
        if (has_p) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float w = __ldg(W + (size_t)c * ld_w + p_col + k);
                s0 = fmaf(pos[3 * (size_t)m0 + k], w, s0);
                s1 = fmaf(pos[3 * (size_t)m1 + k], w, s1);
            }
        }
        const float t0 = tanhf(s0), t1 = tanhf(s1);
        delta[3 * (size_t)m0 + c] = t0;
        new_pos[3 * (size_t)m0 + c] = pos[3 * (size_t)m0 + c] + t0;
        if (m0 + 1 < n) {
            delta[3 * (size_t)m1 + c] = t1;
            new_pos[3 * (size_t)m1 + c] = pos[3 * (size_t)m1 + c] + t1;
        }
    }
}

// gpre = g * (1 - delta^2);  gx = gpre Wx (n x Kx);  gpos = g + gpre Wp (n x 3);  gpre is also written (n x 3, for dW)
__global__ void __launch_bounds__(256) k_head_bwd(const float* __restrict__ g, const float* __restrict__ delta,
                                                  const float* __restrict__ W, int ld_w, int x_col, int p_col, int has_p, int n,
                                                  int Kx, float* __restrict__ gpre, float* __restrict__ gx, int ld_gx,
                                                  float* __restrict__ gpos) {
    const int m = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (m >= n) return;
    float e[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float t = delta[3 * (size_t)m + c];
        e[c] = g[3 * (size_t)m + c] * (1.f - t * t);
    }
    if (lane_id() < 3) {
        const int c = lane_id();
        gpre[3 * (size_t)m + c] = c == 0 ? e[0] : (c == 1 ? e[1] : e[2]);
        if (gpos) {
            float s = g[3 * (size_t)m + c];
            if (has_p)
                s += e[0] * __ldg(W + p_col + c) + e[1] * __ldg(W + ld_w + p_col + c) + e[2] * __ldg(W + 2 * (size_t)ld_w + p_col + c);
            gpos[3 * (size_t)m + c] = s;
        }
    }
    if (gx)
        for (int k = lane_id(); k < Kx; k += 32)
            gx[(size_t)m * ld_gx + k] = e[0] * __ldg(W + x_col + k) + e[1] * __ldg(W + ld_w + x_col + k) +
                                        e[2] * __ldg(W + 2 * (size_t)ld_w + x_col + k);
}

}  // namespace gc2
}  // namespace mrb

using namespace mrb;
using namespace mrb::gc2;

extern "C" int mrb_vert_align_texrows(const float* pos, const int32_t* vert_mesh, const int32_t* mesh_info, int SV, int map_size,
                                      int32_t* texrow, void* stream_) {
    MRB_REQUIRE(pos && vert_mesh && mesh_info && texrow && map_size > 0, "vert_align_texrows: bad arguments");
    if (SV == 0) return MRB_OK;
    k_texrows<<<ceil_div(SV, 256), 256, 0, (cudaStream_t)stream_>>>(pos, vert_mesh, mesh_info, SV, map_size, texrow);
    return check_launch("vert_align_texrows");
}

extern "C" int mrb_gc_gather_fwd(const int32_t* rowptr, const int32_t* col, int n, int D, const float* y, int ld_y,
                                 const float* pos, const float* wp0, const float* wp1, const int32_t* texrow, const float* T,
                                 int relu, uint32_t* mask, const float* residual, int ld_res, float* out, int ld_out,
                                 void* stream_) {
    MRB_REQUIRE(rowptr && col && out && n >= 0 && D > 0 && D % 4 == 0, "gc_gather_fwd: bad arguments");
    MRB_REQUIRE(y || pos || texrow, "gc_gather_fwd: no input term");
    MRB_REQUIRE(!pos || (wp0 && wp1), "gc_gather_fwd: position weights missing");
    MRB_REQUIRE(!texrow || T, "gc_gather_fwd: texel projections missing");
    auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
    MRB_REQUIRE(ld_out % 4 == 0 && al16(out) && (!y || (ld_y % 4 == 0 && al16(y) && ld_y >= 2 * D)) && (!T || al16(T)) &&
                    (!pos || (al16(wp0) && al16(wp1))) && (!residual || (ld_res % 4 == 0 && al16(residual))),
                "gc_gather_fwd: rows must be 16-byte aligned");
    if (n == 0) return MRB_OK;
    FwdParams p;
    p.rowptr = rowptr; p.col = col; p.n = n; p.D = D; p.y = y; p.ld_y = ld_y; p.pos = pos; p.wp0 = wp0; p.wp1 = wp1;
    p.texrow = texrow; p.T = T; p.relu = relu; p.mask = mask; p.residual = residual; p.ld_res = ld_res; p.out = out;
    p.ld_out = ld_out;
    const dim3 grid(min(ceil_div(n, 8), kNumSMs * GC2_FWD_MINB * GC2_WAVES));
    cudaStream_t s = (cudaStream_t)stream_;
    const int sel = (y ? 4 : 0) | (pos ? 2 : 0) | (texrow ? 1 : 0);
    switch (sel) {
        case 1: k_gather_fwd<false, false, true><<<grid, 256, 0, s>>>(p); break;
        case 2: k_gather_fwd<false, true, false><<<grid, 256, 0, s>>>(p); break;
        case 3: k_gather_fwd<false, true, true><<<grid, 256, 0, s>>>(p); break;
        case 4: k_gather_fwd<true, false, false><<<grid, 256, 0, s>>>(p); break;
        case 5: k_gather_fwd<true, false, true><<<grid, 256, 0, s>>>(p); break;
        case 6: k_gather_fwd<true, true, false><<<grid, 256, 0, s>>>(p); break;
        default: k_gather_fwd<true, true, true><<<grid, 256, 0, s>>>(p); break;
    }
    return check_launch("gc_gather_fwd");
}

extern "C" int mrb_gc_gather_bwd(const int32_t* rowptr_t, const int32_t* col_t, int n, int D, const float* gout, int ld_g,
                                 const uint32_t* mask, float* gy, const float* wp0, const float* wp1, float* gpos,
                                 const int32_t* texrow, float* gT, long long tex_rows, void* stream_) {
    MRB_REQUIRE(rowptr_t && col_t && gout && gy && n >= 0 && D > 0 && D % 4 == 0 && ld_g >= D, "gc_gather_bwd: bad arguments");
    MRB_REQUIRE(!gpos || (wp0 && wp1), "gc_gather_bwd: position weights missing");
    MRB_REQUIRE((texrow == nullptr) == (gT == nullptr), "gc_gather_bwd: texrow and gT go together");
    MRB_REQUIRE(((uintptr_t)gy & 15) == 0 && (!gT || ((uintptr_t)gT & 15) == 0), "gc_gather_bwd: gy / gT must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream_;
    if (gT && tex_rows > 0) {
        cudaError_t e = cudaMemsetAsync(gT, 0, sizeof(float) * (size_t)tex_rows * 2 * D, s);      // gT is overwritten
        if (e != cudaSuccess) {
            set_error("gc_gather_bwd: memset: %s", cudaGetErrorString(e));
            return MRB_ERR_CUDA;
        }
    }
    if (n == 0) return MRB_OK;
    BwdParams p;
    p.rowptr_t = rowptr_t; p.col_t = col_t; p.n = n; p.D = D; p.gout = gout; p.ld_g = ld_g; p.mask = mask; p.gy = gy;
    p.wp0 = wp0; p.wp1 = wp1; p.gpos = gpos; p.texrow = texrow; p.gT = gT;
    const bool vec = (ld_g % 4 == 0) && (((uintptr_t)gout & 15) == 0);
    const dim3 grid(min(ceil_div(n, 8), kNumSMs * GC2_BWD_MINB * GC2_WAVES));
    if (gpos) {
        if (vec) k_gather_bwd<true, true><<<grid, 256, 0, s>>>(p);
        else k_gather_bwd<false, true><<<grid, 256, 0, s>>>(p);
    } else {
        if (vec) k_gather_bwd<true, false><<<grid, 256, 0, s>>>(p);
        else k_gather_bwd<false, false><<<grid, 256, 0, s>>>(p);
    }
    return check_launch("gc_gather_bwd");
}

extern "C" int mrb_head_fwd(const float* x, int ld_x, int Kx, const float* pos, const float* W, int ld_w, int x_col, int p_col,
                            int n, float* new_pos, float* delta, void* stream_) {
    MRB_REQUIRE(x && pos && W && new_pos && delta && Kx > 0 && ld_x >= Kx && x_col >= 0, "head_fwd: bad arguments");
    if (n == 0) return MRB_OK;
    k_head_fwd<<<ceil_div(n, 16), 256, 0, (cudaStream_t)stream_>>>(x, ld_x, Kx, pos, W, ld_w, x_col, p_col < 0 ? 0 : p_col,
                                                                  p_col >= 0, n, new_pos, delta);
    return check_launch("head_fwd");
}

extern "C" int mrb_head_bwd(const float* g, const float* delta, const float* W, int ld_w, int x_col, int p_col, int n, int Kx,
                            float* gpre, float* gx, int ld_gx, float* gpos, void* stream_) {
    MRB_REQUIRE(g && delta && W && gpre && Kx > 0 && (!gx || ld_gx >= Kx), "head_bwd: bad arguments");
    if (n == 0) return MRB_OK;
    k_head_bwd<<<ceil_div(n, 8), 256, 0, (cudaStream_t)stream_>>>(g, delta, W, ld_w, x_col, p_col < 0 ? 0 : p_col, p_col >= 0, n,
                                                                 Kx, gpre, gx, ld_gx, gpos);
    return check_launch("head_bwd");
}

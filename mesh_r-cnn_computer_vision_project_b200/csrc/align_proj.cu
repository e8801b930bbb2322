// VertexAlign fused with the linear layer that follows it, and the bf16 feature-map mode.
//
// In the ShapeNet refinement stages VertexAlign (reference meshRCNN/layers.py:509-613) feeds a bias-free bottleneck
// `linear(3840 -> 128)` (layers.py:115,151-155 residual head; :192,230 plain head).  VertexAlign is a masked row gather
// (out[v, c] = fmap[img, c, x1, y1] * [x2 > x1 and y2 > y1], see vert_align.cu), and a gather commutes with a linear map:
//
//     linear(align(f))[v, :] = sum_m  mask_{v,m} * ( f_m[img(v), :, texel_m(v)] . W_m^T )
//                            = sum_m  mask_{v,m} * T_m[img(v) * HW_m + texel_m(v), :]        T_m = rows(f_m) @ W_m^T
//
// with W_m the column slice of the weight that belongs to map m.  So the n_img * sum_m HW_m texels (1 655 per image for
// the four ResNet50 maps of a 137 x 137 input) are projected ONCE per step on the tensor cores (gemm_tc.cu), and every
// vertex gathers and sums <= 4 rows of D = 128 floats.  The SV x 3840 matrix (3.4 GB at BASELINE configs[2]) is never
// formed and the three big GEMMs over it (forward, input gradient, weight gradient) shrink 4x - 40x.
//
// Kernels here:
//   k_map_to_rows / k_rows_to_map   NCHW (fp32 | bf16) <-> channels-last fp32 rows (the A operand of the texel GEMM)
//   k_proj_gather_fwd               out[v, :] = sum_m mask * T[row_m(v), :]                (one warp per vertex)
//   k_proj_gather_bwd               gT[row_m(v), :] += mask * gout[v, :]                   (red.global.add.v4.f32)
//   k_map_to_rows_bf16 + k_fwd_bf16 bf16 feature-map mode of the plain VertexAlign: channels-last bf16 rows, 16-byte
//                                   loads of 8 channels, fp32 output (north-star tolerance for bf16 features: rtol 2e-2)
#include <cuda_bf16.h>

#include "common.cuh"
#include "valign.cuh"
#include "../../include/meshrcnn_b200.h"

namespace mrb {
namespace valign {

constexpr int MAX_MAPS = 8;

struct MapSet {
    int n_maps;
    int size[MAX_MAPS];             // Hm == Wm of map m
    long long row_base[MAX_MAPS];   // first row of map m in the packed texel matrix: n_img * sum_{m' < m} HW_m'
};

__device__ __forceinline__ float load_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// NCHW -> channels-last fp32 rows: per image a C x HW matrix is transposed to HW x C through a 32 x 33 shared tile.
template <typename T>
__global__ void __launch_bounds__(256) k_map_to_rows(const T* __restrict__ src, float* __restrict__ dst, int C, int HW) {
    __shared__ float t[32][33];
    const int img = blockIdx.z;
    const T* s = src + (size_t)img * C * HW;
    float* d = dst + (size_t)img * C * HW;
    const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, pix = p0 + tx;
        t[r][tx] = (c < C && pix < HW) ? load_as_float(s + (size_t)c * HW + pix) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int pix = p0 + r, c = c0 + tx;
        if (pix < HW && c < C) d[(size_t)pix * C + c] = t[tx][r];
    }
}

// channels-last fp32 rows -> NCHW fp32 (gradient of the maps)
__global__ void __launch_bounds__(256) k_rows_to_map(const float* __restrict__ src, int ld_src, float* __restrict__ dst, int C,
                                                     int HW) {
    __shared__ float t[32][33];
    const int img = blockIdx.z;
    const float* s = src + (size_t)img * HW * ld_src;
    float* d = dst + (size_t)img * C * HW;
    const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int pix = p0 + r, c = c0 + tx;
        t[r][tx] = (c < C && pix < HW) ? s[(size_t)pix * ld_src + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, pix = p0 + tx;
        if (pix < HW && c < C) d[(size_t)c * HW + pix] = t[tx][r];
    }
}

// NCHW bf16 -> channels-last bf16 rows
__global__ void __launch_bounds__(256) k_map_to_rows_bf16(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                          int C, int HW) {
    __shared__ __nv_bfloat16 t[32][34];
    const int img = blockIdx.z;
    const __nv_bfloat16* s = src + (size_t)img * C * HW;
    __nv_bfloat16* d = dst + (size_t)img * C * HW;
    const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, pix = p0 + tx;
        t[r][tx] = (c < C && pix < HW) ? s[(size_t)c * HW + pix] : __float2bfloat16(0.f);
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int pix = p0 + r, c = c0 + tx;
        if (pix < HW && c < C) d[(size_t)pix * C + c] = t[tx][r];
    }
}

// texel row of vertex v in map m of the packed texel matrix, or -1 when the vertex is masked in that map
__device__ __forceinline__ long long texel_row(const MapSet& ms, int m, const float* __restrict__ pos,
                                               const int32_t* __restrict__ vert_mesh, const int32_t* __restrict__ mesh_info,
                                               int v) {
    const int sz = ms.size[m];
    const Texel t = project(pos, vert_mesh, mesh_info, v, sz, sz);
    return t.valid ? ms.row_base[m] + (long long)t.img * sz * sz + t.xy : -1;
}

// out[v, :D] = sum_m mask_{v,m} T[row_m(v), :D].  One warp per vertex: lane m < n_maps projects into map m, the rows are
// broadcast, and every lane owns 4 of each 128 columns; the <= 8 row loads of a vertex are independent (issued before the
// adds).  D % 4 == 0, T rows and out rows 16-byte aligned.
__global__ void __launch_bounds__(256) k_proj_gather_fwd(const float* __restrict__ T, int D, const __grid_constant__ MapSet ms,
                                                         const float* __restrict__ pos, const int32_t* __restrict__ vert_mesh,
                                                         const int32_t* __restrict__ mesh_info, int SV,
                                                         float* __restrict__ out, int ld_out) {
    const int v = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (v >= SV) return;
    long long mine = -1;
    if (lane_id() < ms.n_maps) mine = texel_row(ms, lane_id(), pos, vert_mesh, mesh_info, v);
    long long rows[MAX_MAPS];
#pragma unroll
    for (int m = 0; m < MAX_MAPS; ++m) rows[m] = __shfl_sync(0xffffffffu, mine, m);
    for (int d = lane_id() * 4; d < D; d += 128) {
        float4 x[MAX_MAPS];
#pragma unroll
        for (int m = 0; m < MAX_MAPS; ++m) {
            x[m] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < ms.n_maps && rows[m] >= 0) x[m] = __ldg(reinterpret_cast<const float4*>(T + (size_t)rows[m] * D + d));
        }
        float4 acc = x[0];
#pragma unroll
        for (int m = 1; m < MAX_MAPS; ++m) { acc.x += x[m].x; acc.y += x[m].y; acc.z += x[m].z; acc.w += x[m].w; }
        *reinterpret_cast<float4*>(out + (size_t)v * ld_out + d) = acc;
    }
}

// gT[row_m(v), :] += gout[v, :] for every unmasked (vertex, map); vector fp32 reductions (the texel rows are shared by the
// vertices that project to the same texel).  Fully masked vertices return before touching gout.
__global__ void __launch_bounds__(256) k_proj_gather_bwd(const float* __restrict__ gout, int ld_g, int D, const __grid_constant__ MapSet ms,
                                                         const float* __restrict__ pos, const int32_t* __restrict__ vert_mesh,
                                                         const int32_t* __restrict__ mesh_info, int SV, float* __restrict__ gT) {
    const int v = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (v >= SV) return;
    long long mine = -1;
    if (lane_id() < ms.n_maps) mine = texel_row(ms, lane_id(), pos, vert_mesh, mesh_info, v);
    if (__ballot_sync(0xffffffffu, mine >= 0) == 0) return;
    long long rows[MAX_MAPS];
#pragma unroll
    for (int m = 0; m < MAX_MAPS; ++m) rows[m] = __shfl_sync(0xffffffffu, mine, m);
    const bool vec = ((ld_g & 3) == 0) && ((reinterpret_cast<uintptr_t>(gout) & 15) == 0);
    for (int d = lane_id() * 4; d < D; d += 128) {
        float4 g;
        const float* gp = gout + (size_t)v * ld_g + d;
        if (vec) g = __ldg(reinterpret_cast<const float4*>(gp));
        else g = make_float4(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2), __ldg(gp + 3));
#pragma unroll
        for (int m = 0; m < MAX_MAPS; ++m)
            if (m < ms.n_maps && rows[m] >= 0)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gT + (size_t)rows[m] * D + d), "f"(g.x),
                             "f"(g.y), "f"(g.z), "f"(g.w)
                             : "memory");
    }
}

// bf16 feature-map mode of the plain VertexAlign: a warp owns VPW vertices per iteration (their row loads are all issued
// before the first conversion), lane = 8 consecutive channels of a channels-last bf16 row (one 16-byte load), fp32 output.
constexpr int BF_VPW = 4;
__global__ void __launch_bounds__(256) k_fwd_bf16(const __nv_bfloat16* __restrict__ rows_cl, int C, int Hm, int Wm,
                                                  const float* __restrict__ pos, const int32_t* __restrict__ vert_mesh,
                                                  const int32_t* __restrict__ mesh_info, int SV, float* __restrict__ out,
                                                  int ld_out) {
    const int v0 = (blockIdx.x * (blockDim.x >> 5) + warp_id()) * BF_VPW;
    if (v0 >= SV) return;
    long long mine = -1;
    if (lane_id() < BF_VPW && v0 + lane_id() < SV) {
        const Texel t = project(pos, vert_mesh, mesh_info, v0 + lane_id(), Hm, Wm);
        if (t.valid) mine = (long long)t.img * Hm * Wm + t.xy;
    }
    long long rows[BF_VPW];
#pragma unroll
    for (int i = 0; i < BF_VPW; ++i) rows[i] = __shfl_sync(0xffffffffu, mine, i);
    for (int c = lane_id() * 8; c < C; c += 256) {
        uint4 raw[BF_VPW];
#pragma unroll
        for (int i = 0; i < BF_VPW; ++i) {
            raw[i] = make_uint4(0u, 0u, 0u, 0u);
            if (rows[i] >= 0) raw[i] = __ldg(reinterpret_cast<const uint4*>(rows_cl + (size_t)rows[i] * C + c));
        }
#pragma unroll
        for (int i = 0; i < BF_VPW; ++i) {
            if (v0 + i >= SV) break;
            const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
            float f[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {            // bf16 -> fp32 is a 16-bit shift
                f[2 * q] = __uint_as_float(w[q] << 16);
                f[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
            }
            float* dst = out + (size_t)(v0 + i) * ld_out + c;
            *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(f[4], f[5], f[6], f[7]);
        }
    }
}

// unaligned / odd channel counts: lane per channel straight from the NCHW bf16 map
__global__ void __launch_bounds__(256) k_fwd_bf16_scalar(const __nv_bfloat16* __restrict__ fmap, int C, int Hm, int Wm,
                                                         const float* __restrict__ pos, const int32_t* __restrict__ vert_mesh,
                                                         const int32_t* __restrict__ mesh_info, int SV, float* __restrict__ out,
                                                         int ld_out) {
    const int v = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (v >= SV) return;
    const Texel t = project(pos, vert_mesh, mesh_info, v, Hm, Wm);
    const size_t plane = (size_t)Hm * Wm;
    const __nv_bfloat16* src = fmap + (size_t)t.img * C * plane + t.xy;
    float* dst = out + (size_t)v * ld_out;
    for (int c = lane_id(); c < C; c += 32) dst[c] = t.valid ? __bfloat162float(src[(size_t)c * plane]) : 0.f;
}

static int make_mapset(MapSet& ms, int n_maps, const int* map_size_host, int n_img, const char* what) {
    MRB_REQUIRE(n_maps >= 1 && n_maps <= MAX_MAPS && map_size_host, "%s: 1..%d feature maps", what, MAX_MAPS);
    ms.n_maps = n_maps;
    long long base = 0;
    for (int m = 0; m < MAX_MAPS; ++m) {
        ms.size[m] = m < n_maps ? map_size_host[m] : 1;
        ms.row_base[m] = base;
        if (m < n_maps) {
            MRB_REQUIRE(map_size_host[m] > 0, "%s: bad map size", what);
            base += (long long)n_img * map_size_host[m] * map_size_host[m];
        }
    }
    return MRB_OK;
}

}  // namespace valign
}  // namespace mrb

using namespace mrb;
using namespace mrb::valign;

extern "C" int mrb_feature_map_to_rows(const void* fmap, int dtype, int n_img, int C, int HW, float* rows, void* stream_) {
    MRB_REQUIRE(fmap && rows && n_img >= 0 && C > 0 && HW > 0, "feature_map_to_rows: bad arguments");
    MRB_REQUIRE(dtype == 0 || dtype == 1, "feature_map_to_rows: dtype must be 0 (fp32) or 1 (bf16)");
    MRB_REQUIRE(n_img <= 65535, "feature_map_to_rows: too many images");
    if (n_img == 0) return MRB_OK;
    const dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), n_img);
    cudaStream_t s = (cudaStream_t)stream_;
    if (dtype == 0) k_map_to_rows<float><<<grid, 256, 0, s>>>((const float*)fmap, rows, C, HW);
    else k_map_to_rows<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)fmap, rows, C, HW);
    return check_launch("feature_map_to_rows");
}

extern "C" int mrb_rows_to_feature_map(const float* rows, int ld_rows, int n_img, int C, int HW, float* gfmap, void* stream_) {
    MRB_REQUIRE(rows && gfmap && n_img >= 0 && C > 0 && HW > 0 && ld_rows >= C, "rows_to_feature_map: bad arguments");
    MRB_REQUIRE(n_img <= 65535, "rows_to_feature_map: too many images");
    if (n_img == 0) return MRB_OK;
    k_rows_to_map<<<dim3(ceil_div(HW, 32), ceil_div(C, 32), n_img), 256, 0, (cudaStream_t)stream_>>>(rows, ld_rows, gfmap, C, HW);
    return check_launch("rows_to_feature_map");
}

extern "C" int mrb_vert_align_proj_fwd(const float* T, int D, int n_maps, const int* map_size_host, int n_img,
                                       const float* pos, const int32_t* vert_mesh, const int32_t* mesh_info, int SV,
                                       float* out, int ld_out, void* stream_) {
    MRB_REQUIRE(T && pos && vert_mesh && mesh_info && out, "vert_align_proj_fwd: null pointer");
    MRB_REQUIRE(D > 0 && D % 4 == 0 && ld_out % 4 == 0 && ld_out >= D && ((uintptr_t)T & 15) == 0 && ((uintptr_t)out & 15) == 0,
                "vert_align_proj_fwd: D and the output rows must be 16-byte aligned");
    MapSet ms;
    if (int rc = make_mapset(ms, n_maps, map_size_host, n_img, "vert_align_proj_fwd")) return rc;
    if (SV == 0) return MRB_OK;
    k_proj_gather_fwd<<<ceil_div(SV, 8), 256, 0, (cudaStream_t)stream_>>>(T, D, ms, pos, vert_mesh, mesh_info, SV, out, ld_out);
    return check_launch("vert_align_proj_fwd");
}

extern "C" int mrb_vert_align_proj_bwd(const float* gout, int ld_g, int D, int n_maps, const int* map_size_host, int n_img,
                                       const float* pos, const int32_t* vert_mesh, const int32_t* mesh_info, int SV, float* gT,
                                       void* stream_) {
    MRB_REQUIRE(gout && pos && vert_mesh && mesh_info && gT, "vert_align_proj_bwd: null pointer");
    MRB_REQUIRE(D > 0 && D % 4 == 0 && ld_g >= D && ((uintptr_t)gT & 15) == 0, "vert_align_proj_bwd: D % 4 and 16-byte aligned gT");
    MapSet ms;
    if (int rc = make_mapset(ms, n_maps, map_size_host, n_img, "vert_align_proj_bwd")) return rc;
    cudaStream_t s = (cudaStream_t)stream_;
    long long total_rows = 0;
    for (int m = 0; m < n_maps; ++m) total_rows += (long long)n_img * map_size_host[m] * map_size_host[m];
    cudaError_t e = cudaMemsetAsync(gT, 0, sizeof(float) * (size_t)total_rows * D, s);     // gT is overwritten, not accumulated
    if (e != cudaSuccess) {
        set_error("vert_align_proj_bwd: memset: %s", cudaGetErrorString(e));
        return MRB_ERR_CUDA;
    }
    if (SV == 0) return MRB_OK;
    k_proj_gather_bwd<<<ceil_div(SV, 8), 256, 0, s>>>(gout, ld_g, D, ms, pos, vert_mesh, mesh_info, SV, gT);
    return check_launch("vert_align_proj_bwd");
}

extern "C" int mrb_vert_align_fwd_bf16(const void* fmap, int n_img, int C, int Hm, int Wm, const float* pos,
                                       const int32_t* vert_mesh, const int32_t* mesh_info, int SV, float* out, int ld_out,
                                       void* workspace, void* stream_) {
    MRB_REQUIRE(fmap && pos && vert_mesh && mesh_info && out, "vert_align_fwd_bf16: null pointer");
    MRB_REQUIRE(Hm == Wm, "vert_align: feature maps must be square (the reference indexes H with the x coordinate)");
    MRB_REQUIRE(n_img <= 65535, "vert_align_fwd_bf16: too many images");
    if (SV == 0) return MRB_OK;
    cudaStream_t s = (cudaStream_t)stream_;
    const __nv_bfloat16* f = (const __nv_bfloat16*)fmap;
    const bool vec = workspace && (C % 8 == 0) && (ld_out % 4 == 0) && (((uintptr_t)out & 15) == 0) &&
                     (((uintptr_t)workspace & 15) == 0);
    if (!vec) {
        k_fwd_bf16_scalar<<<ceil_div(SV, 8), 256, 0, s>>>(f, C, Hm, Wm, pos, vert_mesh, mesh_info, SV, out, ld_out);
        return check_launch("vert_align_fwd_bf16");
    }
    const int HW = Hm * Wm;
    k_map_to_rows_bf16<<<dim3(ceil_div(HW, 32), ceil_div(C, 32), n_img), 256, 0, s>>>(f, (__nv_bfloat16*)workspace, C, HW);
    k_fwd_bf16<<<ceil_div(SV, 8 * BF_VPW), 256, 0, s>>>((const __nv_bfloat16*)workspace, C, Hm, Wm, pos, vert_mesh, mesh_info,
                                                        SV, out, ld_out);
    return check_launch("vert_align_fwd_bf16");
}

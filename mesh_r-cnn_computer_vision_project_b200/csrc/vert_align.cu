// VertexAlign: camera projection of mesh vertices + feature-map gather, and its scatter-add backward.
//
// Reproduces the *actual* arithmetic of reference VertexAlign (meshRCNN/layers.py:521-613), which is not
// bilinear: the `.long()` cast at :592 turns the four bilinear weights into integers, so
//     out[v, c] = fmap[img(v), c, x1, y1] * [x2 > x1 and y2 > y1]
// with x1 = floor(x), x2 = min(ceil(x), size_x-1) derived from the *w* pixel coordinate but indexing the H axis
// of the map, and y from *h* indexing the W axis (:587-590).  No gradient reaches the vertex positions.
// Index arithmetic is done with explicitly rounded fp32 ops (no FMA contraction, true division) so that the
// floor boundaries match the reference's separate torch ops bit for bit.
//
// Layout: maps stay NCHW fp32 (the reference layout); they are small (<= 2.5 MB per image) and L2 resident, so
// the channel-strided reads are L2 hits while the SV x C output -- the only HBM-heavy stream -- is written with
// fully coalesced 128-bit stores (one warp per vertex, lanes over channels).
#include "common.cuh"
#include "../../include/meshrcnn_b200.h"

namespace mrb {
namespace valign {

struct Texel {
    int img;     // image index
    int xy;      // x1 * Wm + y1
    int valid;   // mask
};

// per-mesh record: image index, image height, image width
__device__ __forceinline__ Texel project(const float* __restrict__ pos, const int32_t* __restrict__ vert_mesh,
                                         const int32_t* __restrict__ mesh_info, int v, int Hm, int Wm) {
    const int mesh = vert_mesh[v];
    const int img = mesh_info[3 * mesh + 0];
    const float H = (float)mesh_info[3 * mesh + 1], W = (float)mesh_info[3 * mesh + 2];
    const float p0 = pos[3 * (size_t)v + 0], p1 = pos[3 * (size_t)v + 1], p2 = pos[3 * (size_t)v + 2];
    // layers.py:557-558 (separate fp32 ops: div, mul, add)
    float h = __fadd_rn(__fmul_rn(248.f, __fdiv_rn(p1, p2)), 111.5f);
    float w = __fadd_rn(__fmul_rn(248.f, __fdiv_rn(p0, -p2)), 111.5f);
    // :561-562  clamp(min=0, max=H-1)
    h = fminf(fmaxf(h, 0.f), H - 1.f);
    w = fminf(fmaxf(w, 0.f), W - 1.f);
    // :577-578  divisor is a python double cast to fp32; true division
    const float sx = (float)((double)mesh_info[3 * mesh + 2] / (double)Wm);
    const float sy = (float)((double)mesh_info[3 * mesh + 1] / (double)Hm);
    const float x = __fdiv_rn(w, sx), y = __fdiv_rn(h, sy);
    const int x1 = (int)floorf(x), y1 = (int)floorf(y);
    const int x2 = min((int)ceilf(x), Wm - 1), y2 = min((int)ceilf(y), Hm - 1);   // :583-584
    Texel t;
    t.img = img;
    // x (from w, scaled by size_x = last dim) indexes the H axis; y indexes the W axis (:587)
    t.xy = x1 * Wm + y1;
    t.valid = (x2 > x1) && (y2 > y1) && x1 >= 0 && y1 >= 0 && x1 < Hm && y1 < Wm;
    return t;
}

__global__ void __launch_bounds__(256) k_fwd(const float* __restrict__ fmap, int C, int Hm, int Wm,
                                             const float* __restrict__ pos, const int32_t* __restrict__ vert_mesh,
                                             const int32_t* __restrict__ mesh_info, int SV, float* __restrict__ out,
                                             int ld_out) {
    const int v = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (v >= SV) return;
    const Texel t = project(pos, vert_mesh, mesh_info, v, Hm, Wm);
    const size_t plane = (size_t)Hm * Wm;
    const float* src = fmap + (size_t)t.img * C * plane + t.xy;
    float* dst = out + (size_t)v * ld_out;
    for (int c = lane_id(); c < C; c += 32) dst[c] = t.valid ? __ldg(src + (size_t)c * plane) : 0.f;
}

__global__ void __launch_bounds__(256) k_bwd(const float* __restrict__ gout, int ld_g, int C, int Hm, int Wm,
                                             const float* __restrict__ pos, const int32_t* __restrict__ vert_mesh,
                                             const int32_t* __restrict__ mesh_info, int SV, float* __restrict__ gfmap) {
    const int v = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (v >= SV) return;
    const Texel t = project(pos, vert_mesh, mesh_info, v, Hm, Wm);
    if (!t.valid) return;
    const size_t plane = (size_t)Hm * Wm;
    float* dst = gfmap + (size_t)t.img * C * plane + t.xy;
    const float* g = gout + (size_t)v * ld_g;
    for (int c = lane_id(); c < C; c += 32) atomicAdd(dst + (size_t)c * plane, g[c]);
}

}  // namespace valign
}  // namespace mrb

using namespace mrb;
using namespace mrb::valign;

extern "C" int mrb_vert_align_fwd(const float* fmap, int n_img, int C, int Hm, int Wm, const float* pos,
                                  const int32_t* vert_mesh, const int32_t* mesh_info, int SV, float* out, int ld_out,
                                  void* stream_) {
    MRB_REQUIRE(fmap && pos && vert_mesh && mesh_info && out, "vert_align_fwd: null pointer");
    MRB_REQUIRE(Hm == Wm, "vert_align: feature maps must be square (the reference indexes H with the x coordinate)");
    (void)n_img;
    if (SV == 0) return MRB_OK;
    k_fwd<<<ceil_div(SV, 8), 256, 0, (cudaStream_t)stream_>>>(fmap, C, Hm, Wm, pos, vert_mesh, mesh_info, SV, out, ld_out);
    return check_launch("vert_align_fwd");
}

extern "C" int mrb_vert_align_bwd(const float* gout, int ld_g, int n_img, int C, int Hm, int Wm, const float* pos,
                                  const int32_t* vert_mesh, const int32_t* mesh_info, int SV, float* gfmap,
                                  void* stream_) {
    MRB_REQUIRE(gout && pos && vert_mesh && mesh_info && gfmap, "vert_align_bwd: null pointer");
    (void)n_img;
    if (SV == 0) return MRB_OK;
    k_bwd<<<ceil_div(SV, 8), 256, 0, (cudaStream_t)stream_>>>(gout, ld_g, C, Hm, Wm, pos, vert_mesh, mesh_info, SV, gfmap);
    return check_launch("vert_align_bwd");
}

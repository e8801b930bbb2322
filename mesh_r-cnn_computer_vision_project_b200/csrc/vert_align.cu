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
// Layout / data movement (forward): the maps arrive NCHW fp32 (the reference layout) and are small (<= 2.5 MB per image,
// L2 resident); the SV x C output is the HBM-heavy stream (3.4 GB at config 3).  The maps are first transposed to
// channels-last (one pass over ~80 MB), so that the C channels of a texel are one contiguous row; the gather itself is
// then done entirely by the TMA engine: every lane owns one vertex and issues a bulk copy (cp.async.bulk, UBLKCP) of
// its texel row global -> shared, waits on the warp's mbarrier, and issues the bulk store shared -> global of the
// output row (masked vertices store a row of zeros).  No feature value passes through a register, loads and stores are
// full 128-byte lines, and 16-32 rows per warp are in flight.  Rows that are not 16-byte aligned (C % 4 != 0) fall back
// to the lane-per-channel kernel below.
#include "common.cuh"
#include "valign.cuh"
#include "../../include/meshrcnn_b200.h"

namespace mrb {
namespace valign {

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

// NCHW -> NHWC: per image a C x HW matrix is transposed to HW x C through a 32 x 33 shared tile.
__global__ void __launch_bounds__(256) k_to_channels_last(const float* __restrict__ src, float* __restrict__ dst, int C, int HW) {
    __shared__ float t[32][33];
    const int img = blockIdx.z;
    const float* s = src + (size_t)img * C * HW;
    float* d = dst + (size_t)img * C * HW;
    const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, pix = p0 + tx;
        t[r][tx] = (c < C && pix < HW) ? s[(size_t)c * HW + pix] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int pix = p0 + r, c = c0 + tx;
        if (pix < HW && c < C) d[(size_t)pix * C + c] = t[tx][r];
    }
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// TMA gather: grid-stride over vertex batches; warp = `slots` vertices per batch, lane = vertex.
constexpr int VA_WARPS = 4;
constexpr int VA_WARP_BYTES = 16384;

__global__ void __launch_bounds__(VA_WARPS * 32) k_fwd_bulk(const float* __restrict__ fmap_cl, int C, int Hm, int Wm,
                                                           const float* __restrict__ pos, const int32_t* __restrict__ vert_mesh,
                                                           const int32_t* __restrict__ mesh_info, int SV,
                                                           float* __restrict__ out, int ld_out, int slots) {
    extern __shared__ __align__(128) unsigned char va_smem[];
    __shared__ __align__(8) unsigned long long bars[VA_WARPS];
    const int row_bytes = C * 4;
    unsigned char* zero_row = va_smem;                                         // row_bytes of zeros (shared by the CTA)
    unsigned char* my_slots = va_smem + ((row_bytes + 127) & ~127) + warp_id() * VA_WARP_BYTES;
    for (int i = threadIdx.x; i < C; i += blockDim.x) reinterpret_cast<float*>(zero_row)[i] = 0.f;
    const uint32_t bar = smem_addr(&bars[warp_id()]);
    if (lane_id() == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                // zero row + barrier visible to the async proxy
    __syncthreads();

    const size_t plane = (size_t)Hm * Wm;
    uint32_t phase = 0;
    const int batches = (SV + slots - 1) / slots;
    for (int b = blockIdx.x * VA_WARPS + warp_id(); b < batches; b += gridDim.x * VA_WARPS) {
        const int v = b * slots + lane_id();
        const bool mine = lane_id() < slots && v < SV;
        Texel t;
        t.valid = 0;
        if (mine) t = project(pos, vert_mesh, mesh_info, v, Hm, Wm);
        const unsigned valid_mask = __ballot_sync(0xffffffffu, mine && t.valid);
        const uint32_t slot = smem_addr(my_slots + lane_id() * row_bytes);
        if (valid_mask) {
            if (lane_id() == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                             "r"((uint32_t)(__popc(valid_mask) * row_bytes))
                             : "memory");
            __syncwarp();
            if (mine && t.valid) {
                const float* src = fmap_cl + ((size_t)t.img * plane + t.xy) * C;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(slot),
                             "l"(src), "r"((uint32_t)row_bytes), "r"(bar)
                             : "memory");
            }
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "VA_WAIT:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra.uni VA_DONE;\n\t"
                "bra.uni VA_WAIT;\n\t"
                "VA_DONE:\n\t}" ::"r"(bar), "r"(phase) : "memory");
            phase ^= 1;
        }
        if (mine) {
            float* dst = out + (size_t)v * ld_out;
            const uint32_t srcs = t.valid ? slot : smem_addr(zero_row);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(srcs),
                         "r"((uint32_t)row_bytes)
                         : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");        // slots may be overwritten by the next batch
        __syncwarp();
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
                                  float* workspace, void* stream_) {
    MRB_REQUIRE(fmap && pos && vert_mesh && mesh_info && out, "vert_align_fwd: null pointer");
    MRB_REQUIRE(Hm == Wm, "vert_align: feature maps must be square (the reference indexes H with the x coordinate)");
    if (SV == 0) return MRB_OK;
    cudaStream_t s = (cudaStream_t)stream_;
    const bool bulk = workspace && (C % 4 == 0) && (ld_out % 4 == 0) && (((uintptr_t)out & 15) == 0) &&
                      (((uintptr_t)workspace & 15) == 0) && C * 4 <= VA_WARP_BYTES;
    if (!bulk) {
        k_fwd<<<ceil_div(SV, 8), 256, 0, s>>>(fmap, C, Hm, Wm, pos, vert_mesh, mesh_info, SV, out, ld_out);
        return check_launch("vert_align_fwd");
    }
    const int HW = Hm * Wm;
    k_to_channels_last<<<dim3(ceil_div(HW, 32), ceil_div(C, 32), n_img), 256, 0, s>>>(fmap, workspace, C, HW);
    const int row_bytes = C * 4;
    const int slots = min(32, VA_WARP_BYTES / row_bytes);
    const size_t smem = ((row_bytes + 127) & ~127) + (size_t)VA_WARPS * VA_WARP_BYTES;
    static SmemOptIn optin;
    if (int rc = ensure_dynamic_smem(k_fwd_bulk, 16384 + VA_WARPS * VA_WARP_BYTES, optin, "vert_align_fwd")) return rc;
    const int batches = ceil_div(SV, slots);
    const int grid = min(ceil_div(batches, VA_WARPS), 2 * kNumSMs);
    k_fwd_bulk<<<grid, VA_WARPS * 32, smem, s>>>(workspace, C, Hm, Wm, pos, vert_mesh, mesh_info, SV, out, ld_out, slots);
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

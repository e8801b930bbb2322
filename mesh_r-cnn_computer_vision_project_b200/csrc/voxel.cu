// Voxel-head tail: the step immediately before Cubify in the reference models.
//
//   voxelGrid = VoxelBranch(...)            ends in nn.Sigmoid           (meshRCNN/layers.py:487-506)
//   voxel_loss = BCE(voxelGrid, gt, mean)                                (meshRCNN/loss_functions.py:10-14, shapenet_model.py:64)
//   cubify(voxelGrid)                       thresholds the probabilities (shapenet_model.py:70-71)
//
// Evaluated eagerly that is three passes over the B x V^3 grid plus the probability tensor's round trip through HBM.
// Here the sigmoid is folded into its two consumers: mrb_voxel_bce_fwd reads the LOGITS once and produces the mean BCE (and
// the probabilities only if the caller asks for them), and mrb_cubify_count thresholds sigmoid(logit) in its first kernel
// (cubify.cu, `from_logits`).  BCE keeps torch's semantics: log terms clamped at -100, fp64 accumulation, mean over all
// elements; the backward is (p - t) / max(p (1 - p), 1e-12) w.r.t. probabilities and (p - t) w.r.t. logits.
#include "common.cuh"
#include "../../include/meshrcnn_b200.h"

namespace mrb {
namespace voxel {

__device__ __forceinline__ float sigmoid_f32(float x) { return 1.0f / (1.0f + expf(-x)); }   // torch's CUDA formula, full-precision expf

template <bool FROM_LOGITS>
__global__ void __launch_bounds__(256) k_bce_fwd(const float* __restrict__ x, const float* __restrict__ t, long long n,
                                                 float* __restrict__ probs, double* __restrict__ acc) {
    __shared__ double sd[33];
    double a = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float p = FROM_LOGITS ? sigmoid_f32(x[i]) : x[i];
        if (FROM_LOGITS && probs) probs[i] = p;
        const float ti = t[i];
        const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(log1pf(-p), -100.f);   // torch clamps both logs at -100
        a -= (double)(ti * lp + (1.f - ti) * lq);
    }
    a = block_sum<double>(a, sd);
    if (threadIdx.x == 0) atomicAdd(acc, a);
}

__global__ void k_mean(const double* __restrict__ acc, double inv_n, float* __restrict__ out) {
    if (threadIdx.x == 0) out[0] = (float)(acc[0] * inv_n);
}

template <bool FROM_LOGITS>
__global__ void __launch_bounds__(256) k_bce_bwd(const float* __restrict__ x, const float* __restrict__ t, long long n,
                                                 const float* __restrict__ g, float inv_n, float* __restrict__ gx) {
    const float s = (*g) * inv_n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float ti = t[i];
        if (FROM_LOGITS) {
            gx[i] = s * (sigmoid_f32(x[i]) - ti);
        } else {
            const float p = x[i];
            gx[i] = s * (p - ti) / fmaxf(p * (1.f - p), 1e-12f);
        }
    }
}

}  // namespace voxel
}  // namespace mrb

using namespace mrb;
using namespace mrb::voxel;

extern "C" int mrb_voxel_bce_fwd(const float* x, const float* target, long long n, int from_logits, float* probs_out,
                                 double* acc, float* loss_out, void* stream_) {
    MRB_REQUIRE(x && target && acc && loss_out && n > 0, "voxel_bce_fwd: bad arguments");
    cudaStream_t s = (cudaStream_t)stream_;
    cudaMemsetAsync(acc, 0, sizeof(double), s);
    const int blocks = (int)min((long long)8 * kNumSMs, ceil_div64(n, 256));
    if (from_logits) k_bce_fwd<true><<<blocks, 256, 0, s>>>(x, target, n, probs_out, acc);
    else k_bce_fwd<false><<<blocks, 256, 0, s>>>(x, target, n, nullptr, acc);
    k_mean<<<1, 32, 0, s>>>(acc, 1.0 / (double)n, loss_out);
    return check_launch("voxel_bce_fwd");
}

extern "C" int mrb_voxel_bce_bwd(const float* x, const float* target, long long n, int from_logits, const float* g,
                                 float* gx, void* stream_) {
    MRB_REQUIRE(x && target && g && gx && n > 0, "voxel_bce_bwd: bad arguments");
    const int blocks = (int)min((long long)8 * kNumSMs, ceil_div64(n, 256));
    if (from_logits) k_bce_bwd<true><<<blocks, 256, 0, (cudaStream_t)stream_>>>(x, target, n, g, (float)(1.0 / (double)n), gx);
    else k_bce_bwd<false><<<blocks, 256, 0, (cudaStream_t)stream_>>>(x, target, n, g, (float)(1.0 / (double)n), gx);
    return check_launch("voxel_bce_bwd");
}

// Shared helpers for the meshrcnn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#define MRB_OK 0
#define MRB_ERR_ARG 1
#define MRB_ERR_CUDA 2
#define MRB_ERR_EMPTY 3

namespace mrb {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return MRB_ERR_CUDA;
    }
    return MRB_OK;
}

#define MRB_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            mrb::set_error(__VA_ARGS__);  \
            return MRB_ERR_ARG;           \
        }                                 \
    } while (0)

// Opt-in to > 48 KB of dynamic shared memory.  cudaFuncSetAttribute applies to the *current device*, so the "already
// done" state is one bit per device ordinal (a process may drive several GPUs from several threads, like the reference's
// thread-per-GPU CustomDP, dataParallel/dataParallel.py:33).  The attribute call is idempotent, so two threads racing on
// the same device both succeed; the bit is only a fast path.  Devices >= 64 set the attribute on every call.
struct SmemOptIn {
    std::atomic<unsigned long long> done{0};
};
template <typename KernelT>
inline int ensure_dynamic_smem(KernelT kernel, int bytes, SmemOptIn& st, const char* what) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("%s: cudaGetDevice: %s", what, cudaGetErrorString(e));
        return MRB_ERR_CUDA;
    }
    const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (bit && (st.done.load(std::memory_order_acquire) & bit)) return MRB_OK;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        set_error("%s: cannot reserve %d bytes of dynamic shared memory on device %d: %s", what, bytes, dev,
                  cudaGetErrorString(e));
        return MRB_ERR_CUDA;
    }
    if (bit) st.done.fetch_or(bit, std::memory_order_release);
    return MRB_OK;
}

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline long long ceil_div64(long long a, long long b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------------------
// warp / block primitives
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane_id() >= o) v += n;
    }
    return v;
}

// Block-wide sum; result valid in every thread. `scratch` must hold >= 32 elements of T.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
    v = warp_sum(v);
    __syncthreads();
    if (lane_id() == 0) scratch[warp_id()] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    T t = (threadIdx.x < nw) ? scratch[threadIdx.x] : T(0);
    if (warp_id() == 0) {
        t = warp_sum(t);
        if (lane_id() == 0) scratch[0] = t;
    }
    __syncthreads();
    return scratch[0];
}

// Block-wide exclusive scan of one int per thread. Returns the exclusive prefix; *total gets the block sum.
// `scratch` must hold >= 33 ints.
__device__ __forceinline__ int block_exclusive_scan(int v, int* scratch, int* total) {
    int inc = warp_inclusive_scan(v);
    __syncthreads();
    if (lane_id() == 31) scratch[warp_id()] = inc;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    if (warp_id() == 0) {
        int w = (lane_id() < nw) ? scratch[lane_id()] : 0;
        int winc = warp_inclusive_scan(w);
        scratch[lane_id()] = winc - w;
        if (lane_id() == 31) scratch[32] = winc;
    }
    __syncthreads();
    int base = scratch[warp_id()];
    *total = scratch[32];
    return base + inc - v;
}

__device__ __forceinline__ void atomic_add_f64(double* addr, double v) { atomicAdd(addr, v); }

}  // namespace mrb

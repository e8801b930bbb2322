// Point-cloud "normals" by local PCA, the normal-consistency loss, the edge-length loss, and their backward.
//
// Replaces compute_normals / batched_normal_distance (reference meshRCNN/loss_functions.py:107-170: gather,
// mean, centre, 3x3 scatter matrix, `S.cpu()` -> LAPACK symeig -> `.to(device)`, argmin, gather, dot, abs, sum)
// and total_edge_length over a dense SVxSV matrix (:47-48,175-189).  Everything stays on the device; the 3x3
// symmetric eigenproblem is solved per point with cyclic Jacobi rotations in fp64.
//
// Reference semantics kept on purpose (see DESIGN.md "normal loss"):
//   * the k-NN indices of point i of cloud X index the *other* cloud, but are used to gather rows of X (:141,146);
//   * the "normal" is *row* argmin(eigenvalues) (= row 0, eigenvalues ascending) of the eigenvector matrix V,
//     i.e. (v0[0], v1[0], v2[0]), because `eigen_vectors[b, p, argmin]` indexes the row dimension (:165-168).
// The only freedom is the sign of each eigenvector (LAPACK's choice is unspecified); this kernel fixes it by:
// V[2][0] >= 0;  largest-|.| component of column 1 positive;  det V = +1.
#include "common.cuh"
#include "../../include/meshrcnn_b200.h"

namespace mrb {
namespace normals {

constexpr int KMAX = 16;

// Eigen-decomposition of a symmetric 3x3 (a = [xx, xy, xz, yy, yz, zz]); eigenvalues ascending in w,
// eigenvectors in the columns of V, canonical signs applied.
__device__ void eigh3(const double a[6], double w[3], double V[3][3]) {
    double A[3][3] = {{a[0], a[1], a[2]}, {a[1], a[3], a[4]}, {a[2], a[4], a[5]}};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    const double scale = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]) + 1e-300;
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off <= 1e-17 * scale) break;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int p = (r == 2) ? 1 : 0, q = (r == 0) ? 1 : 2;   // (0,1), (0,2), (1,2)
            const double apq = A[p][q];
            if (fabs(apq) <= 1e-300) continue;
            const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            const int o = 3 - p - q;
            const double app = A[p][p], aqq = A[q][q];
            A[p][p] = app - t * apq;
            A[q][q] = aqq + t * apq;
            A[p][q] = A[q][p] = 0.0;
            const double aop = A[o][p], aoq = A[o][q];
            A[o][p] = A[p][o] = c * aop - s * aoq;
            A[o][q] = A[q][o] = s * aop + c * aoq;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double vip = V[i][p], viq = V[i][q];
                V[i][p] = c * vip - s * viq;
                V[i][q] = s * vip + c * viq;
            }
        }
    }
    w[0] = A[0][0]; w[1] = A[1][1]; w[2] = A[2][2];
    // sort ascending (3-element network), swapping eigenvector columns
#define MRB_SWAP(i, j)                                                       \
    if (w[i] > w[j]) {                                                       \
        double tw = w[i]; w[i] = w[j]; w[j] = tw;                            \
        for (int r_ = 0; r_ < 3; ++r_) { double tv = V[r_][i]; V[r_][i] = V[r_][j]; V[r_][j] = tv; } \
    }
    MRB_SWAP(0, 1) MRB_SWAP(1, 2) MRB_SWAP(0, 1)
#undef MRB_SWAP
    // canonical signs
    if (V[2][0] < 0.0) for (int r = 0; r < 3; ++r) V[r][0] = -V[r][0];
    {
        double vb = V[0][1];                                    // largest-|.| component of column 1 (first wins ties)
        if (fabs(V[1][1]) > fabs(vb)) vb = V[1][1];
        if (fabs(V[2][1]) > fabs(vb)) vb = V[2][1];
        if (vb < 0.0) for (int r = 0; r < 3; ++r) V[r][1] = -V[r][1];
    }
    const double det = V[0][0] * (V[1][1] * V[2][2] - V[1][2] * V[2][1]) - V[0][1] * (V[1][0] * V[2][2] - V[1][2] * V[2][0]) +
                       V[0][2] * (V[1][0] * V[2][1] - V[1][1] * V[2][0]);
    if (det < 0.0) for (int r = 0; r < 3; ++r) V[r][2] = -V[r][2];
}

// gathers the k neighbours of point (batch, p) from pt, returns centred rows Y and the scatter matrix.
// KT > 0: k is the compile-time constant KT, the loops unroll and Y stays in registers (with a run-time k the dynamically
// indexed Y lives in local memory and the kernel is bound by it); KT = 0: generic k <= KMAX.
template <int KT>
__device__ __forceinline__ void neighbourhood(const float* __restrict__ pt, const int32_t* __restrict__ nn, int k_rt,
                                              double Y[KMAX][3], double S[6]) {
    const int k = KT > 0 ? KT : k_rt;
    double m[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < (KT > 0 ? KT : KMAX); ++j) {
        if (j >= k) break;
        const float* r = pt + 3 * (size_t)nn[j];
        Y[j][0] = r[0]; Y[j][1] = r[1]; Y[j][2] = r[2];
        m[0] += Y[j][0]; m[1] += Y[j][1]; m[2] += Y[j][2];
    }
    m[0] /= k; m[1] /= k; m[2] /= k;
    for (int c = 0; c < 6; ++c) S[c] = 0.0;
#pragma unroll
    for (int j = 0; j < (KT > 0 ? KT : KMAX); ++j) {
        if (j >= k) break;
        const double y0 = Y[j][0] - m[0], y1 = Y[j][1] - m[1], y2 = Y[j][2] - m[2];
        Y[j][0] = y0; Y[j][1] = y1; Y[j][2] = y2;
        S[0] += y0 * y0; S[1] += y0 * y1; S[2] += y0 * y2; S[3] += y1 * y1; S[4] += y1 * y2; S[5] += y2 * y2;
    }
}

// resident blocks per SM asked of ptxas for the backward kernel (scripts/variants.sh normals.cu "python scripts/time_normals.py"):
// 1 (145 registers): 67.4 us; 5 (96): 59.4; 6 (80, 48 B spills): 59.4; 8 (64): 65.5; 10 (48): 80.5.  The forward kernel is flat
// (43 - 44 us for 1 .. 8).
#ifndef MRB_NORMALS_MINB
#define MRB_NORMALS_MINB 5
#endif
template <int KT>
__global__ void __launch_bounds__(128) k_normals_fwd(const float* __restrict__ pt, const int32_t* __restrict__ knn, int P,
                                                     int k, float* __restrict__ normals, double* __restrict__ eig,
                                                     size_t eig_plane) {
    const int batch = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const size_t o = (size_t)batch * P + p;
    double Y[KMAX][3], S[6], w[3], V[3][3];
    neighbourhood<KT>(pt + (size_t)batch * P * 3, knn + o * k, k, Y, S);
    eigh3(S, w, V);
    normals[3 * o] = (float)V[0][0];
    normals[3 * o + 1] = (float)V[0][1];
    normals[3 * o + 2] = (float)V[0][2];
    if (eig) {      // 12 planes of B * P doubles (w0 w1 w2 | V row-major): the backward pass skips the Jacobi sweeps
#pragma unroll
        for (int c = 0; c < 3; ++c) eig[c * eig_plane + o] = w[c];
#pragma unroll
        for (int c = 0; c < 9; ++c) eig[(3 + c) * eig_plane + o] = V[c / 3][c % 3];
    }
}

// gn -> gpt (atomic scatter to the gathered rows):  n_j = V[0][j]
//   K_ij = V[0][i] gn_j / (w_j - w_i) (i != j),  gS = V K V^T,  gY = Y (gS + gS^T)
// PAD4: gpt rows are 4 floats wide (xyz + one unused lane), so a neighbour's three partial sums go out as ONE 16-byte
// red.global.add.v4.f32 instead of three scalar atomics (the kernel is bound by the L2 atomic units: 30 -> 10 per point).
template <int KT, bool PAD4>
__global__ void __launch_bounds__(128, MRB_NORMALS_MINB) k_normals_bwd(const float* __restrict__ pt, const int32_t* __restrict__ knn, int P,
                                                     int k, const float* __restrict__ gn, float* __restrict__ gpt,
                                                     const double* __restrict__ eig, size_t eig_plane) {
    const int batch = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const size_t o = (size_t)batch * P + p;
    const double g[3] = {gn[3 * o], gn[3 * o + 1], gn[3 * o + 2]};
    if (g[0] == 0.0 && g[1] == 0.0 && g[2] == 0.0) return;
    double Y[KMAX][3], S[6], w[3], V[3][3];
    const int32_t* nn = knn + o * k;
    neighbourhood<KT>(pt + (size_t)batch * P * 3, nn, k, Y, S);
    if (eig) {      // eigen-decomposition saved by the forward pass
#pragma unroll
        for (int c = 0; c < 3; ++c) w[c] = eig[c * eig_plane + o];
#pragma unroll
        for (int c = 0; c < 9; ++c) V[c / 3][c % 3] = eig[(3 + c) * eig_plane + o];
    } else {
        eigh3(S, w, V);
    }
    double Kf[3][3];
    const double tiny = 1e-14 * (fabs(w[0]) + fabs(w[1]) + fabs(w[2])) + 1e-300;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const double gap = w[j] - w[i];
            Kf[i][j] = (i == j || fabs(gap) < tiny) ? 0.0 : V[0][i] * g[j] / gap;
        }
    double T[3][3], G[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[i][j] = V[i][0] * Kf[0][j] + V[i][1] * Kf[1][j] + V[i][2] * Kf[2][j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) G[i][j] = T[i][0] * V[j][0] + T[i][1] * V[j][1] + T[i][2] * V[j][2];
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) { const double s = G[i][j] + G[j][i]; G[i][j] = s; G[j][i] = s; }
    float* gb = gpt + (size_t)batch * P * (PAD4 ? 4 : 3);
#pragma unroll
    for (int j = 0; j < (KT > 0 ? KT : KMAX); ++j) {
        if (j >= (KT > 0 ? KT : k)) break;
        const float gx = (float)(Y[j][0] * G[0][0] + Y[j][1] * G[1][0] + Y[j][2] * G[2][0]);
        const float gy = (float)(Y[j][0] * G[0][1] + Y[j][1] * G[1][1] + Y[j][2] * G[2][1]);
        const float gz = (float)(Y[j][0] * G[0][2] + Y[j][1] * G[1][2] + Y[j][2] * G[2][2]);
        if (PAD4) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gb + 4 * (size_t)nn[j]), "f"(gx), "f"(gy), "f"(gz),
                         "f"(0.f)
                         : "memory");
        } else {
            const size_t r = 3 * (size_t)nn[j];
            atomicAdd(gb + r, gx);
            atomicAdd(gb + r + 1, gy);
            atomicAdd(gb + r + 2, gz);
        }
    }
}

// acc[0] += sum_i |n_a[i] . n_b[idx_a[i]]| ;  acc[1] += sum_j |n_b[j] . n_a[idx_b[j]]|   (loss_functions.py:119-125)
__global__ void __launch_bounds__(256) k_normal_loss(const float* __restrict__ na, const float* __restrict__ nb, int P, int Q,
                                                     const int32_t* __restrict__ idx_a, const int32_t* __restrict__ idx_b,
                                                     double* __restrict__ acc) {
    __shared__ double sd[33];
    const int batch = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double l0 = 0.0, l1 = 0.0;
    if (t < P) {
        const size_t i = (size_t)batch * P + t, j = (size_t)batch * Q + idx_a[i];
        l0 = fabs((double)na[3 * i] * nb[3 * j] + (double)na[3 * i + 1] * nb[3 * j + 1] + (double)na[3 * i + 2] * nb[3 * j + 2]);
    }
    if (t < Q) {
        const size_t j = (size_t)batch * Q + t, i = (size_t)batch * P + idx_b[j];
        l1 = fabs((double)na[3 * i] * nb[3 * j] + (double)na[3 * i + 1] * nb[3 * j + 1] + (double)na[3 * i + 2] * nb[3 * j + 2]);
    }
    l0 = block_sum<double>(l0, sd);
    __syncthreads();
    l1 = block_sum<double>(l1, sd);
    if (threadIdx.x == 0) { atomicAdd(acc, l0); atomicAdd(acc + 1, l1); }
}

__global__ void __launch_bounds__(256) k_normal_loss_bwd(const float* __restrict__ na, const float* __restrict__ nb, int P,
                                                         int Q, const int32_t* __restrict__ idx_a,
                                                         const int32_t* __restrict__ idx_b, const float* __restrict__ g0,
                                                         const float* __restrict__ g1, float scale, float* __restrict__ gna,
                                                         float* __restrict__ gnb) {
    const int batch = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int pass = 0; pass < 2; ++pass) {
        size_t i, j;
        float g;
        if (pass == 0) {
            if (t >= P) continue;
            i = (size_t)batch * P + t; j = (size_t)batch * Q + idx_a[i]; g = scale * (*g0);
        } else {
            if (t >= Q) continue;
            j = (size_t)batch * Q + t; i = (size_t)batch * P + idx_b[j]; g = scale * (*g1);
        }
        const float dot = na[3 * i] * nb[3 * j] + na[3 * i + 1] * nb[3 * j + 1] + na[3 * i + 2] * nb[3 * j + 2];
        const float sg = dot > 0.f ? g : (dot < 0.f ? -g : 0.f);
        for (int d = 0; d < 3; ++d) {
            if (gna) atomicAdd(gna + 3 * i + d, sg * nb[3 * j + d]);
            if (gnb) atomicAdd(gnb + 3 * j + d, sg * na[3 * i + d]);
        }
    }
}

__global__ void k_finalize2(const double* __restrict__ acc, int n, double scale, float* __restrict__ out) {
    if (threadIdx.x < n) out[threadIdx.x] = (float)(acc[threadIdx.x] * scale);
}
__global__ void k_finalize_sum2(const double* __restrict__ acc, double scale, float* __restrict__ out) {
    if (threadIdx.x == 0) out[0] = (float)((acc[0] + acc[1]) * scale);
}

// small scalar glue of the loss (sums over stages, the weighted total): one launch instead of a chain of ATen kernels
constexpr int SC_MAX = 16;
struct ScalarSet { const float* x[SC_MAX]; float w[SC_MAX]; int n; };
__global__ void k_scalar_combine(const __grid_constant__ ScalarSet s, float* __restrict__ out) {
    if (threadIdx.x == 0) {
        float acc = 0.f;
        for (int i = 0; i < s.n; ++i) acc = fmaf(s.w[i], *s.x[i], acc);
        out[0] = acc;
    }
}
__global__ void k_scalar_scatter(const float* __restrict__ g, const __grid_constant__ ScalarSet s, float* __restrict__ out) {
    if (threadIdx.x < s.n) out[threadIdx.x] = (*g) * s.w[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------------------
// edge loss: mean over the directed edge list of |v_r - v_c|^2   (O(E), no SVxSV matrix)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_edge_fwd(const float* __restrict__ pos, const long long* __restrict__ rows,
                                                  const long long* __restrict__ cols, long long E,
                                                  double* __restrict__ acc) {
    __shared__ double sd[33];
    double a = 0.0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
        const long long r = rows[e], c = cols[e];
        const float dx = pos[3 * r] - pos[3 * c], dy = pos[3 * r + 1] - pos[3 * c + 1], dz = pos[3 * r + 2] - pos[3 * c + 2];
        a += (double)dx * dx + (double)dy * dy + (double)dz * dz;
    }
    a = block_sum<double>(a, sd);
    if (threadIdx.x == 0) atomicAdd(acc, a);
}

__global__ void __launch_bounds__(256) k_edge_bwd(const float* __restrict__ pos, const long long* __restrict__ rows,
                                                  const long long* __restrict__ cols, long long E,
                                                  const float* __restrict__ g, float* __restrict__ gpos) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const float k = 2.f * (*g) / (float)E;
    const long long r = rows[e], c = cols[e];
    for (int d = 0; d < 3; ++d) {
        const float v = k * (pos[3 * r + d] - pos[3 * c + d]);
        atomicAdd(gpos + 3 * r + d, v);
        atomicAdd(gpos + 3 * c + d, -v);
    }
}

}  // namespace normals
}  // namespace mrb

using namespace mrb;
using namespace mrb::normals;

static int launch_normals_fwd(const float* pt, const int32_t* knn, int B, int P, int k, float* normals_out, double* eig,
                              void* stream_) {
    MRB_REQUIRE(pt && knn && normals_out, "normals_fwd: null pointer");
    MRB_REQUIRE(k >= 1 && k <= KMAX, "normals_fwd: k must be in [1, %d]", KMAX);
    if (B == 0 || P == 0) return MRB_OK;
    const dim3 grid(ceil_div(P, 128), B);
    const size_t plane = (size_t)B * P;
    if (k == 10) k_normals_fwd<10><<<grid, 128, 0, (cudaStream_t)stream_>>>(pt, knn, P, k, normals_out, eig, plane);   // the reference default
    else k_normals_fwd<0><<<grid, 128, 0, (cudaStream_t)stream_>>>(pt, knn, P, k, normals_out, eig, plane);
    return check_launch("normals_fwd");
}

extern "C" int mrb_normals_fwd(const float* pt, const int32_t* knn, int B, int P, int k, float* normals_out, void* stream_) {
    return launch_normals_fwd(pt, knn, B, P, k, normals_out, nullptr, stream_);
}

extern "C" int mrb_normals_fwd_eig(const float* pt, const int32_t* knn, int B, int P, int k, float* normals_out, double* eig,
                                   void* stream_) {
    return launch_normals_fwd(pt, knn, B, P, k, normals_out, eig, stream_);
}

static int launch_normals_bwd(const float* pt, const int32_t* knn, int B, int P, int k, const float* gn, float* gpt, int ld_gpt,
                              const double* eig, void* stream_) {
    MRB_REQUIRE(pt && knn && gn && gpt, "normals_bwd: null pointer");
    MRB_REQUIRE(k >= 1 && k <= KMAX, "normals_bwd: k must be in [1, %d]", KMAX);
    MRB_REQUIRE(ld_gpt == 3 || (ld_gpt == 4 && ((uintptr_t)gpt & 15) == 0),
                "normals_bwd: gradient rows must be 3 floats, or 4 floats and 16-byte aligned (got ld = %d)", ld_gpt);
    if (B == 0 || P == 0) return MRB_OK;
    const dim3 grid(ceil_div(P, 128), B);
    cudaStream_t s = (cudaStream_t)stream_;
    const size_t plane = (size_t)B * P;
    if (ld_gpt == 4) {
        if (k == 10) k_normals_bwd<10, true><<<grid, 128, 0, s>>>(pt, knn, P, k, gn, gpt, eig, plane);
        else k_normals_bwd<0, true><<<grid, 128, 0, s>>>(pt, knn, P, k, gn, gpt, eig, plane);
    } else {
        if (k == 10) k_normals_bwd<10, false><<<grid, 128, 0, s>>>(pt, knn, P, k, gn, gpt, eig, plane);
        else k_normals_bwd<0, false><<<grid, 128, 0, s>>>(pt, knn, P, k, gn, gpt, eig, plane);
    }
    return check_launch("normals_bwd");
}

extern "C" int mrb_normals_bwd(const float* pt, const int32_t* knn, int B, int P, int k, const float* gn, float* gpt,
                               void* stream_) {
    return launch_normals_bwd(pt, knn, B, P, k, gn, gpt, 3, nullptr, stream_);
}

extern "C" int mrb_normals_bwd_ld(const float* pt, const int32_t* knn, int B, int P, int k, const float* gn, float* gpt,
                                  int ld_gpt, const double* eig, void* stream_) {
    return launch_normals_bwd(pt, knn, B, P, k, gn, gpt, ld_gpt, eig, stream_);
}

extern "C" int mrb_normal_loss_fwd(const float* na, const float* nb, int B, int P, int Q, const int32_t* idx_a,
                                   const int32_t* idx_b, double* acc2, float* out2, void* stream_) {
    MRB_REQUIRE(na && nb && idx_a && idx_b && acc2 && out2, "normal_loss_fwd: null pointer");
    cudaStream_t s = (cudaStream_t)stream_;
    cudaMemsetAsync(acc2, 0, 2 * sizeof(double), s);
    if (B > 0 && max(P, Q) > 0)
        k_normal_loss<<<dim3(ceil_div(max(P, Q), 256), B), 256, 0, s>>>(na, nb, P, Q, idx_a, idx_b, acc2);
    k_finalize2<<<1, 32, 0, s>>>(acc2, 2, 1.0, out2);
    return check_launch("normal_loss_fwd");
}

extern "C" int mrb_normal_loss_bwd(const float* na, const float* nb, int B, int P, int Q, const int32_t* idx_a,
                                   const int32_t* idx_b, const float* g0, const float* g1, float* gna, float* gnb,
                                   void* stream_) {
    MRB_REQUIRE(na && nb && idx_a && idx_b && g0 && g1, "normal_loss_bwd: null pointer");
    if (B == 0 || max(P, Q) == 0 || (!gna && !gnb)) return MRB_OK;
    k_normal_loss_bwd<<<dim3(ceil_div(max(P, Q), 256), B), 256, 0, (cudaStream_t)stream_>>>(na, nb, P, Q, idx_a, idx_b, g0,
                                                                                           g1, 1.f, gna, gnb);
    return check_launch("normal_loss_bwd");
}

extern "C" int mrb_normal_loss_total_fwd(const float* na, const float* nb, int B, int P, int Q, const int32_t* idx_a,
                                         const int32_t* idx_b, double scale, double* acc2, float* out1, void* stream_) {
    MRB_REQUIRE(na && nb && idx_a && idx_b && acc2 && out1, "normal_loss_total_fwd: null pointer");
    cudaStream_t s = (cudaStream_t)stream_;
    cudaMemsetAsync(acc2, 0, 2 * sizeof(double), s);
    if (B > 0 && max(P, Q) > 0)
        k_normal_loss<<<dim3(ceil_div(max(P, Q), 256), B), 256, 0, s>>>(na, nb, P, Q, idx_a, idx_b, acc2);
    k_finalize_sum2<<<1, 32, 0, s>>>(acc2, scale, out1);
    return check_launch("normal_loss_total_fwd");
}

extern "C" int mrb_normal_loss_total_bwd(const float* na, const float* nb, int B, int P, int Q, const int32_t* idx_a,
                                         const int32_t* idx_b, const float* g, float scale, float* gna, float* gnb,
                                         void* stream_) {
    MRB_REQUIRE(na && nb && idx_a && idx_b && g, "normal_loss_total_bwd: null pointer");
    if (B == 0 || max(P, Q) == 0 || (!gna && !gnb)) return MRB_OK;
    k_normal_loss_bwd<<<dim3(ceil_div(max(P, Q), 256), B), 256, 0, (cudaStream_t)stream_>>>(na, nb, P, Q, idx_a, idx_b, g, g,
                                                                                           scale, gna, gnb);
    return check_launch("normal_loss_total_bwd");
}

static int fill_scalar_set(ScalarSet& s, const float* const* xs_host, const float* w_host, int n, const char* what) {
    MRB_REQUIRE(n >= 1 && n <= SC_MAX && w_host, "%s: 1..%d scalars", what, SC_MAX);
    s.n = n;
    for (int i = 0; i < SC_MAX; ++i) {
        s.x[i] = (i < n && xs_host) ? xs_host[i] : nullptr;
        s.w[i] = i < n ? w_host[i] : 0.f;
    }
    return MRB_OK;
}

extern "C" int mrb_scalar_combine(const void* xs_host, const float* w_host, int n, float* out, void* stream_) {
    MRB_REQUIRE(xs_host && out, "scalar_combine: null pointer");
    ScalarSet s;
    if (int rc = fill_scalar_set(s, (const float* const*)xs_host, w_host, n, "scalar_combine")) return rc;
    for (int i = 0; i < n; ++i) MRB_REQUIRE(s.x[i], "scalar_combine: null scalar %d", i);
    k_scalar_combine<<<1, 32, 0, (cudaStream_t)stream_>>>(s, out);
    return check_launch("scalar_combine");
}

extern "C" int mrb_scalar_scatter(const float* g, const float* w_host, int n, float* out, void* stream_) {
    MRB_REQUIRE(g && out, "scalar_scatter: null pointer");
    ScalarSet s;
    if (int rc = fill_scalar_set(s, nullptr, w_host, n, "scalar_scatter")) return rc;
    k_scalar_scatter<<<1, 32, 0, (cudaStream_t)stream_>>>(g, s, out);
    return check_launch("scalar_scatter");
}

extern "C" int mrb_edge_loss_fwd(const float* pos, const long long* adj, long long E, double* acc, float* out,
                                 void* stream_) {
    MRB_REQUIRE(pos && acc && out && (adj || E == 0), "edge_loss_fwd: null pointer");
    cudaStream_t s = (cudaStream_t)stream_;
    cudaMemsetAsync(acc, 0, sizeof(double), s);
    if (E > 0) {
        const int blocks = (int)min((long long)4 * kNumSMs, ceil_div64(E, 256));
        k_edge_fwd<<<blocks, 256, 0, s>>>(pos, adj, adj + E, E, acc);
    }
    k_finalize2<<<1, 32, 0, s>>>(acc, 1, E > 0 ? 1.0 / (double)E : 0.0, out);
    return check_launch("edge_loss_fwd");
}

extern "C" int mrb_edge_loss_bwd(const float* pos, const long long* adj, long long E, const float* g, float* gpos,
                                 void* stream_) {
    MRB_REQUIRE(pos && g && gpos && (adj || E == 0), "edge_loss_bwd: null pointer");
    if (E == 0) return MRB_OK;
    k_edge_bwd<<<(unsigned)ceil_div64(E, 256), 256, 0, (cudaStream_t)stream_>>>(pos, adj, adj + E, E, g, gpos);
    return check_launch("edge_loss_bwd");
}

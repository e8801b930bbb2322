// Cubify: voxel occupancy probabilities -> packed triangle meshes + sorted COO / CSR adjacency.
//
// Replaces Cubify.forward (reference meshRCNN/layers.py:403-484): threshold + conv3d + nonzero + 6 masked
// index ops + argsort + unique(dim=0) + a Python dict over every vertex/face key + unique(dim=1).
//
// B200 design: every output order of the reference is a scan order:
//   faces     : (b, dir, z, y, x), two triangles per exposed quad
//   vertices  : (b, z, y, x) over the (Z+1)(Y+1)(X+1) corner lattice  (== torch.unique(dim=0) order)
//   adjacency : (row, col)  == CSR with sorted columns                (== torch.unique(dim=1) order)
// so no sort, no hash table and no dictionary are needed: flags -> block counts -> one small scan ->
// (host reads 2B+3 counters once, the API returns Python lists) -> emit.  All kernels are HBM-bound byte /
// integer work: one coalesced pass over the probabilities (4 B/voxel), a 1 B/voxel face-flag array that stays
// L2 resident, and coalesced int64 output streams (24 B/face, 16 B/directed edge, 12 B/vertex).
#include "common.cuh"
#include "../../include/meshrcnn_b200.h"

namespace mrb {
namespace cubify {

constexpr int CH = 512;  // items (voxels or lattice points) per block

// neighbour whose emptiness exposes the face, (dz,dy,dx) per direction -- layers.py:357-362
__constant__ int kNbr[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, -1}, {0, 0, 1}};
// quad corners c0..c3 per direction as lattice offsets {0,1}^3 in (z,y,x) -- layers.py:370-400
// (dirs 2/3 sit on the side opposite the empty neighbour: reference quirk kept for bit-exactness)
__constant__ int kCorner[6][4][3] = {
    {{0, 0, 0}, {0, 0, 1}, {0, 1, 0}, {0, 1, 1}}, {{1, 0, 0}, {1, 0, 1}, {1, 1, 0}, {1, 1, 1}},
    {{1, 0, 0}, {1, 0, 1}, {0, 0, 0}, {0, 0, 1}}, {{0, 1, 0}, {0, 1, 1}, {1, 1, 0}, {1, 1, 1}},
    {{1, 0, 0}, {0, 0, 0}, {1, 1, 0}, {0, 1, 0}}, {{0, 0, 1}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}}};
// For a lattice point that is corner `s` (sz<<2|sy<<1|sx, 1 = +1/2 side) of a voxel with exposed face `d`:
// 27-bit mask over neighbour offsets (dz+1)*9+(dy+1)*3+(dx+1) reached through the quad's triangle edges
// c0c1, c0c2, c1c2, c2c3, c0c3 (layers.py:441-443,469-472); 0 if the quad does not touch that corner.
__constant__ uint32_t kLatMask[8][6] = {
    {0x0034000u, 0x0000000u, 0x0c04000u, 0x0000000u, 0x2400000u, 0x0000000u},
    {0x0009000u, 0x0000000u, 0x0201000u, 0x0000000u, 0x0000000u, 0x2410000u},
    {0x0004c00u, 0x0000000u, 0x0000000u, 0x0c04000u, 0x0480000u, 0x0000000u},
    {0x0001200u, 0x0000000u, 0x0000000u, 0x0201000u, 0x0000000u, 0x0480400u},
    {0x0000000u, 0x0034000u, 0x0004030u, 0x0000000u, 0x0010090u, 0x0000000u},
    {0x0000000u, 0x0009000u, 0x0001008u, 0x0000000u, 0x0000000u, 0x0000090u},
    {0x0000000u, 0x0004c00u, 0x0000000u, 0x0004030u, 0x0000412u, 0x0000000u},
    {0x0000000u, 0x0001200u, 0x0000000u, 0x0001008u, 0x0000000u, 0x0000012u}};

struct Dims {
    int B, Z, Y, X;
    int LZ, LY, LX;
    int nvox;   // Z*Y*X
    int nlat;   // LZ*LY*LX
    int nchF;   // voxel chunks per mesh
    int nchL;   // lattice chunks per mesh
};

static Dims make_dims(int B, int Z, int Y, int X) {
    Dims d;
    d.B = B; d.Z = Z; d.Y = Y; d.X = X;
    d.LZ = Z + 1; d.LY = Y + 1; d.LX = X + 1;
    d.nvox = Z * Y * X;
    d.nlat = d.LZ * d.LY * d.LX;
    d.nchF = (d.nvox + CH - 1) / CH;
    d.nchL = (d.nlat + CH - 1) / CH;
    return d;
}

struct Workspace {
    uint8_t* ff;      // [B*nvox] 6 face-flag bits per voxel
    int32_t* rank;    // [B*nlat] vertex id per used lattice point
    int32_t* faceOff; // [B*6*nchF + 1]
    int32_t* vertOff; // [B*nchL + 1]
    int32_t* edgeOff; // [B*nchL + 1]
};

static size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

static size_t carve(const Dims& d, void* base, Workspace* ws) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += align16(bytes);
        return base ? (void*)((char*)base + o) : (void*)nullptr;
    };
    void* p0 = take((size_t)d.B * d.nvox);
    void* p1 = take((size_t)d.B * d.nlat * 4);
    void* p2 = take(((size_t)d.B * 6 * d.nchF + 1) * 4);
    void* p3 = take(((size_t)d.B * d.nchL + 1) * 4);
    void* p4 = take(((size_t)d.B * d.nchL + 1) * 4);
    if (ws) {
        ws->ff = (uint8_t*)p0; ws->rank = (int32_t*)p1; ws->faceOff = (int32_t*)p2;
        ws->vertOff = (int32_t*)p3; ws->edgeOff = (int32_t*)p4;
    }
    return off;
}

// ---------------------------------------------------------------------------------------------------------
// pass 1a: threshold + exposed-face flags + per-(mesh, dir, chunk) face counts
// ---------------------------------------------------------------------------------------------------------
// occupancy test: probs > th (strict, fp32 -- layers.py:405), or sigmoid(logit) > th when the grid holds the voxel head's
// logits (the sigmoid of VoxelBranch folded into Cubify; same formula as voxel.cu / torch's CUDA sigmoid)
template <bool LOGITS>
__device__ __forceinline__ bool occupied(float v, float th) {
    return (LOGITS ? 1.0f / (1.0f + expf(-v)) : v) > th;
}

template <bool LOGITS>
__global__ void __launch_bounds__(CH) k_faceflags(const float* __restrict__ probs, float th, Dims d, Workspace ws) {
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int v = chunk * CH + threadIdx.x;
    unsigned f = 0;
    if (v < d.nvox) {
        const float* p = probs + (size_t)b * d.nvox;
        if (occupied<LOGITS>(__ldg(p + v), th)) {
            const int x = v % d.X, y = (v / d.X) % d.Y, z = v / (d.X * d.Y);
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const int nz = z + kNbr[k][0], ny = y + kNbr[k][1], nx = x + kNbr[k][2];
                bool occ = false;                      // zero padding (layers.py:411)
                if (nz >= 0 && nz < d.Z && ny >= 0 && ny < d.Y && nx >= 0 && nx < d.X)
                    occ = occupied<LOGITS>(__ldg(p + ((size_t)nz * d.Y + ny) * d.X + nx), th);
                if (!occ) f |= 1u << k;
            }
        }
        ws.ff[(size_t)b * d.nvox + v] = (uint8_t)f;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        int c = __syncthreads_count((f >> k) & 1u);
        if (threadIdx.x == 0) ws.faceOff[((size_t)b * 6 + k) * d.nchF + chunk] = c;
    }
}

// neighbour mask (27 bits) of lattice point (lz,ly,lx) of mesh b; 0 <=> not a vertex
__device__ __forceinline__ uint32_t lattice_mask(const uint8_t* __restrict__ ffb, const Dims& d, int lz, int ly, int lx) {
    uint32_t m = 0;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int vz = lz - 1 + ((a >> 2) & 1), vy = ly - 1 + ((a >> 1) & 1), vx = lx - 1 + (a & 1);
        if (vz < 0 || vz >= d.Z || vy < 0 || vy >= d.Y || vx < 0 || vx >= d.X) continue;
        const unsigned f = ffb[((size_t)vz * d.Y + vy) * d.X + vx];
        if (!f) continue;
        const int s = 7 - a;   // corner position relative to that voxel: (1-az, 1-ay, 1-ax)
#pragma unroll
        for (int k = 0; k < 6; ++k)
            if ((f >> k) & 1u) m |= kLatMask[s][k];
    }
    return m;
}

// pass 1b: per-(mesh, lattice chunk) vertex and directed-edge counts
__global__ void __launch_bounds__(CH) k_lattice_count(Dims d, Workspace ws) {
    __shared__ int scratch[33];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int l = chunk * CH + threadIdx.x;
    uint32_t m = 0;
    if (l < d.nlat) {
        const int lx = l % d.LX, ly = (l / d.LX) % d.LY, lz = l / (d.LX * d.LY);
        m = lattice_mask(ws.ff + (size_t)b * d.nvox, d, lz, ly, lx);
    }
    const int nv = __syncthreads_count(m != 0);
    const int ne = block_sum<int>(__popc(m), scratch);
    if (threadIdx.x == 0) {
        ws.vertOff[(size_t)b * d.nchL + chunk] = nv;
        ws.edgeOff[(size_t)b * d.nchL + chunk] = ne;
    }
}

// pass 1c: exclusive scans of the three count arrays (one block each) + per-mesh totals.
// meta layout (int64): [0]=SV [1]=SF [2]=E [3]=unused, then v_count[B], f_count[B], v_offset[B], f_offset[B]
__global__ void __launch_bounds__(1024) k_scan(Dims d, Workspace ws, long long* __restrict__ meta) {
    __shared__ int scratch[33];
    int32_t* arr;
    int n;
    if (blockIdx.x == 0) { arr = ws.faceOff; n = d.B * 6 * d.nchF; }
    else if (blockIdx.x == 1) { arr = ws.vertOff; n = d.B * d.nchL; }
    else { arr = ws.edgeOff; n = d.B * d.nchL; }
    int carry = 0;
    constexpr int PER = 8;                       // consecutive elements per thread: 8192 per block-scan round
    for (int base = 0; base < n; base += 1024 * PER) {
        const int i0 = base + threadIdx.x * PER;
        int vals[PER], sum = 0;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            vals[u] = (i0 + u < n) ? arr[i0 + u] : 0;
            sum += vals[u];
        }
        int total;
        int run = carry + block_exclusive_scan(sum, scratch, &total);
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            if (i0 + u < n) arr[i0 + u] = run;
            run += vals[u];
        }
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) arr[n] = carry;
    __syncthreads();
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) meta[1] = 2LL * carry;
        for (int b = threadIdx.x; b < d.B; b += 1024) {
            meta[4 + d.B + b] = 2LL * (arr[(b + 1) * 6 * d.nchF] - arr[b * 6 * d.nchF]);
            meta[4 + 3 * d.B + b] = 2LL * arr[b * 6 * d.nchF];
        }
    } else if (blockIdx.x == 1) {
        if (threadIdx.x == 0) meta[0] = carry;
        for (int b = threadIdx.x; b < d.B; b += 1024) {
            meta[4 + b] = arr[(b + 1) * d.nchL] - arr[b * d.nchL];
            meta[4 + 2 * d.B + b] = arr[b * d.nchL];
        }
    } else {
        if (threadIdx.x == 0) { meta[2] = carry; meta[3] = 0; }
    }
}

// ---------------------------------------------------------------------------------------------------------
// pass 2a: vertices, lattice ranks, CSR row pointers
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CH) k_emit_verts(Dims d, Workspace ws, float* __restrict__ verts,
                                                   int32_t* __restrict__ rowptr, int32_t* __restrict__ vert_mesh,
                                                   uint32_t* __restrict__ vmask, int32_t* __restrict__ vlat) {
    __shared__ int scratch[33];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int l = chunk * CH + threadIdx.x;
    uint32_t m = 0;
    int lx = 0, ly = 0, lz = 0;
    if (l < d.nlat) {
        lx = l % d.LX; ly = (l / d.LX) % d.LY; lz = l / (d.LX * d.LY);
        m = lattice_mask(ws.ff + (size_t)b * d.nvox, d, lz, ly, lx);
    }
    int tot;
    const int exv = block_exclusive_scan(m != 0 ? 1 : 0, scratch, &tot);
    __syncthreads();
    const int exe = block_exclusive_scan(__popc(m), scratch, &tot);
    if (m) {
        const int vid = ws.vertOff[(size_t)b * d.nchL + chunk] + exv;
        ws.rank[(size_t)b * d.nlat + l] = vid;
        // (z,y,x) half-integers rotated 90 deg about axis 0: (z, x, -y)  (layers.py:465-467, exact in fp32)
        verts[3 * (size_t)vid + 0] = (float)lz - 0.5f;
        verts[3 * (size_t)vid + 1] = (float)lx - 0.5f;
        verts[3 * (size_t)vid + 2] = -((float)ly - 0.5f);
        rowptr[vid] = ws.edgeOff[(size_t)b * d.nchL + chunk] + exe;
        vert_mesh[vid] = b;
        vmask[vid] = m;
        vlat[vid] = b * d.nlat + l;
    }
    if (b == d.B - 1 && chunk == d.nchL - 1 && threadIdx.x == 0) {
        const int SV = ws.vertOff[(size_t)d.B * d.nchL];
        rowptr[SV] = ws.edgeOff[(size_t)d.B * d.nchL];
    }
}

// pass 2b: adjacency.  One thread per vertex enumerates its neighbours in ascending lattice order (== ascending
// vertex id); the block's edges form one contiguous output range, so they are staged in shared memory and streamed out
// with fully coalesced int64 / int32 stores (the 16 B/edge COO + 4 B/edge CSR streams are the bulk of Cubify's traffic).
constexpr int ADJ_BLOCK = 256;
constexpr int ADJ_CAP = ADJ_BLOCK * 18;   // a lattice vertex has at most 18 neighbours (6 axis + 12 face-diagonal)

__global__ void __launch_bounds__(ADJ_BLOCK) k_emit_adj(Dims d, Workspace ws, int SV, const int32_t* __restrict__ rowptr,
                                                        const uint32_t* __restrict__ vmask, const int32_t* __restrict__ vlat,
                                                        long long* __restrict__ adj_row, long long* __restrict__ adj_col,
                                                        int32_t* __restrict__ col32) {
    __shared__ int32_t s_col[ADJ_CAP];
    __shared__ unsigned char s_row[ADJ_CAP];
    const int v0 = blockIdx.x * ADJ_BLOCK;
    const int vid = v0 + threadIdx.x;
    const int e_begin = rowptr[v0];
    const int e_end = rowptr[min(v0 + ADJ_BLOCK, SV)];
    if (vid < SV) {
        uint32_t m = vmask[vid];
        const int lat = vlat[vid];
        int e = rowptr[vid] - e_begin;
        const int sY = d.LX, sZ = d.LX * d.LY;
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            const int dz = bit / 9 - 1, dy = (bit / 3) % 3 - 1, dx = bit % 3 - 1;
            s_col[e] = ws.rank[lat + dz * sZ + dy * sY + dx];
            s_row[e] = (unsigned char)threadIdx.x;
            ++e;
        }
    }
    __syncthreads();
    const int n = e_end - e_begin;
    for (int e = threadIdx.x; e < n; e += ADJ_BLOCK) {
        const int c = s_col[e];
        adj_row[(size_t)e_begin + e] = v0 + s_row[e];
        adj_col[(size_t)e_begin + e] = c;
        col32[(size_t)e_begin + e] = c;
    }
}

// pass 2c: faces in (b, dir, z, y, x) order, two triangles (c0,c1,c2),(c0,c2,c3) per quad, per-mesh local ids.
// A quad is 6 int64 = 48 contiguous bytes and consecutive flagged lanes write consecutive quads, so every thread stores its
// quad straight from registers as three 16-byte vectors (a warp covers up to 1.5 KB of contiguous output); one barrier for
// the per-warp offsets of all six directions.
__global__ void __launch_bounds__(CH) k_emit_faces(Dims d, Workspace ws, const long long* __restrict__ meta,
                                                   long long* __restrict__ faces) {
    __shared__ int wtot[6][CH / 32];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int v = chunk * CH + threadIdx.x;
    const unsigned f = (v < d.nvox) ? ws.ff[(size_t)b * d.nvox + v] : 0u;
    int pre[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const unsigned bal = __ballot_sync(0xffffffffu, (f >> k) & 1u);
        pre[k] = __popc(bal & ((1u << lane_id()) - 1u));
        if (lane_id() == 0) wtot[k][warp_id()] = __popc(bal);
    }
    __syncthreads();
    if (f == 0u) return;
    const int x = v % d.X, y = (v / d.X) % d.Y, z = v / (d.X * d.Y);
    const int voff = (int)meta[4 + 2 * d.B + b];
    const int32_t* rk = ws.rank + (size_t)b * d.nlat;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        if (!((f >> k) & 1u)) continue;
        int before = 0;
        for (int w = 0; w < warp_id(); ++w) before += wtot[k][w];
        long long c[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int lz = z + kCorner[k][j][0], ly = y + kCorner[k][j][1], lx = x + kCorner[k][j][2];
            c[j] = rk[((size_t)lz * d.LY + ly) * d.LX + lx] - voff;
        }
        const long long q = ws.faceOff[((size_t)b * 6 + k) * d.nchF + chunk] + before + pre[k];
        longlong2* o = reinterpret_cast<longlong2*>(faces + q * 6);          // 48-byte quads: 16-byte aligned
        o[0] = make_longlong2(c[0], c[1]);
        o[1] = make_longlong2(c[2], c[0]);
        o[2] = make_longlong2(c[2], c[3]);
    }
}

}  // namespace cubify
}  // namespace mrb

using namespace mrb;
using namespace mrb::cubify;

extern "C" long long mrb_cubify_workspace_bytes(int B, int Z, int Y, int X) {
    if (B <= 0 || Z <= 0 || Y <= 0 || X <= 0) return -1;
    Dims d = make_dims(B, Z, Y, X);
    return (long long)carve(d, nullptr, nullptr);
}

extern "C" int mrb_cubify_count(const float* probs, int B, int Z, int Y, int X, float threshold, int from_logits,
                                void* workspace, long long* meta, void* stream_) {
    MRB_REQUIRE(probs && workspace && meta, "cubify_count: null pointer");
    MRB_REQUIRE(B > 0 && Z > 0 && Y > 0 && X > 0, "cubify_count: bad grid %dx%dx%dx%d", B, Z, Y, X);
    MRB_REQUIRE((long long)B * (Z + 1) * (Y + 1) * (X + 1) < (1LL << 31), "cubify_count: lattice too large for int32");
    MRB_REQUIRE(B <= 65535, "cubify_count: batch too large");
    cudaStream_t stream = (cudaStream_t)stream_;
    Dims d = make_dims(B, Z, Y, X);
    Workspace ws;
    carve(d, workspace, &ws);
    if (from_logits) k_faceflags<true><<<dim3(d.nchF, B), CH, 0, stream>>>(probs, threshold, d, ws);
    else k_faceflags<false><<<dim3(d.nchF, B), CH, 0, stream>>>(probs, threshold, d, ws);
    k_lattice_count<<<dim3(d.nchL, B), CH, 0, stream>>>(d, ws);
    k_scan<<<3, 1024, 0, stream>>>(d, ws, meta);
    return check_launch("cubify_count");
}

extern "C" int mrb_cubify_emit(int B, int Z, int Y, int X, void* workspace, const long long* meta, long long SV,
                               long long SF, long long E, float* verts, long long* faces, long long* adj,
                               int32_t* rowptr, int32_t* col32, int32_t* vert_mesh, void* vert_aux, void* stream_) {
    MRB_REQUIRE(workspace && meta && verts && faces && adj && rowptr && col32 && vert_mesh && vert_aux,
                "cubify_emit: null pointer");
    MRB_REQUIRE(SV > 0 && SF > 0 && E > 0, "cubify_emit: empty grid");
    cudaStream_t stream = (cudaStream_t)stream_;
    Dims d = make_dims(B, Z, Y, X);
    Workspace ws;
    carve(d, workspace, &ws);
    uint32_t* vmask = (uint32_t*)vert_aux;
    int32_t* vlat = (int32_t*)vert_aux + SV;
    k_emit_verts<<<dim3(d.nchL, B), CH, 0, stream>>>(d, ws, verts, rowptr, vert_mesh, vmask, vlat);
    k_emit_adj<<<ceil_div(SV, ADJ_BLOCK), ADJ_BLOCK, 0, stream>>>(d, ws, (int)SV, rowptr, vmask, vlat, adj, adj + E, col32);
    k_emit_faces<<<dim3(d.nchF, B), CH, 0, stream>>>(d, ws, meta, faces);
    return check_launch("cubify_emit");
}

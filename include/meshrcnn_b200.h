/* meshrcnn_b200 -- C ABI of the B200-native (sm_100a) Mesh R-CNN voxel-to-mesh refinement hot path.
 *
 * The reference (alondj/Mesh_R-CNN_Computer_Vision_project) has no FFI layer: its boundary is the Python module
 * API of meshRCNN/layers.py, meshRCNN/loss_functions.py and utils/mesh_sampling.py.  The Python package
 * `meshrcnn_b200` mirrors that API and calls the entry points below through ctypes from
 * torch.autograd.Function.forward/backward (see INTEGRATION.md for the binding a maintainer adds).
 *
 * Conventions
 *   - every pointer is a raw *device* pointer unless the name ends in _host; PyTorch (the caller) owns and
 *     allocates all inputs, outputs and workspaces -- the library never allocates, frees or caches memory and
 *     keeps no mutable global state;
 *   - `stream` is a cudaStream_t (the caller's current stream); no entry point synchronises;
 *   - return value 0 = success; otherwise an MRB_ERR_* code, with a message from mrb_last_error()
 *     (thread-local);
 *   - indices at the reference API are int64; "32" suffixed buffers are the int32 CSR copies used internally;
 *   - all floating-point buffers are fp32 unless stated otherwise.
 */
#ifndef MESHRCNN_B200_H
#define MESHRCNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRB_VERSION 100

int mrb_version(void);
const char* mrb_last_error(void);
int mrb_device_info(int* sm_major, int* sm_minor, int* num_sms);

/* ------------------------------------------------------------------------------------------------------------
 * Cubify  -- replaces Cubify.forward, reference meshRCNN/layers.py:403-484.
 *
 * Two phases because output sizes are data dependent and the module API returns Python lists
 * (layers.py:445,448,484):
 *   1. mrb_cubify_count  : threshold (strict >, fp32), exposed-face flags, per-block counts, scans.
 *        meta (int64, 4 + 4*B entries): [0]=total vertices, [1]=total faces, [2]=directed edges E, [3]=0,
 *        then v_count[B], f_count[B], v_offset[B], f_offset[B].
 *   2. (caller copies meta to the host -- the only synchronisation -- and allocates the outputs)
 *   3. mrb_cubify_emit   : vertices (b,z,y,x order, coordinates (z, x, -y), layers.py:447,465-467),
 *        faces (b,dir,z,y,x order, 2 per quad, per-mesh local ids, layers.py:441-443,481-483),
 *        adjacency adj[2][E] sorted by (row, col) (layers.py:469-478) + its CSR form rowptr/col32.
 *        vert_mesh[v] = mesh id of vertex v; vert_aux = 8 bytes per vertex of scratch.
 * workspace: mrb_cubify_workspace_bytes(B,Z,Y,X) bytes, shared by both phases (contents must be preserved).
 * An all-empty batch yields meta[1] == 0; the Python layer raises ValueError("empty grid") like layers.py:434-435.
 */
long long mrb_cubify_workspace_bytes(int B, int Z, int Y, int X);
int mrb_cubify_count(const float* probs, int B, int Z, int Y, int X, float threshold, void* workspace,
                     long long* meta, void* stream);
int mrb_cubify_emit(int B, int Z, int Y, int X, void* workspace, const long long* meta, long long SV, long long SF,
                    long long E, float* verts, long long* faces, long long* adj, int32_t* rowptr, int32_t* col32,
                    int32_t* vert_mesh, void* vert_aux, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MESHRCNN_B200_H */

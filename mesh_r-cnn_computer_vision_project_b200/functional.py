"""Host-side operators: thin ``torch.autograd.Function`` wrappers around the C-ABI kernels.

Nothing here computes on the CPU and nothing falls back to PyTorch ops for the hot path: each function
allocates its outputs with torch (device memory only) and launches hand-written sm_100a kernels on the
caller's current stream through ``_lib.call``.
"""
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib
from .topology import MeshTopology, from_coo, lookup, register

# quad corners per direction as lattice offsets in (z, y, x); same table as csrc/cubify.cu (layers.py:370-400)
CUBIFY_CORNERS = (
    ((0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1)), ((1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)),
    ((1, 0, 0), (1, 0, 1), (0, 0, 0), (0, 0, 1)), ((0, 1, 0), (0, 1, 1), (1, 1, 0), (1, 1, 1)),
    ((1, 0, 0), (0, 0, 0), (1, 1, 0), (0, 1, 0)), ((0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)))


def _require_cuda(t: Tensor, what: str) -> None:
    if not isinstance(t, Tensor) or not t.is_cuda:
        raise RuntimeError("meshrcnn_b200.%s: expected a CUDA tensor -- the hot path has no CPU fallback" % what)


# ----------------------------------------------------------------------------------------------------------
# Cubify
# ----------------------------------------------------------------------------------------------------------
@torch.no_grad()
def cubify(t: Tensor, threshold: float):
    """See ``layers.Cubify``.  Returns (verts, v_index, faces, f_index, adj, topology)."""
    _require_cuda(t, "Cubify")
    if t.dim() != 4:
        raise RuntimeError("Cubify expects B x Z x Y x X, got %s" % (tuple(t.shape),))
    B, Z, Y, X = t.shape
    probs = t.detach().to(torch.float32).contiguous()
    dev = probs.device
    with torch.cuda.device(dev):
        lib = _lib.load()
        ws_bytes = lib.mrb_cubify_workspace_bytes(B, Z, Y, X)
        if ws_bytes < 0:
            raise RuntimeError("Cubify: bad grid shape %s" % (tuple(t.shape),))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        meta = torch.empty(4 + 4 * B, dtype=torch.int64, device=dev)
        _lib.call("mrb_cubify_count", _lib.ptr(probs), B, Z, Y, X, float(threshold), _lib.ptr(ws), _lib.ptr(meta))
        meta_h = meta.cpu()                      # the one unavoidable sync: the API returns Python lists
        SV, SF, E = int(meta_h[0]), int(meta_h[1]), int(meta_h[2])
        if SF == 0:
            raise ValueError("empty grid")       # reference layers.py:434-435
        v_counts = meta_h[4:4 + B].tolist()
        f_counts = meta_h[4 + B:4 + 2 * B].tolist()
        last = max(i for i, c in enumerate(f_counts) if c > 0) + 1   # bincount truncation (layers.py:445,448)
        verts = torch.empty(SV, 3, dtype=torch.float32, device=dev)
        faces = torch.empty(SF, 3, dtype=torch.int64, device=dev)
        adj = torch.empty(2, E, dtype=torch.int64, device=dev)
        rowptr = torch.empty(SV + 1, dtype=torch.int32, device=dev)
        col32 = torch.empty(E, dtype=torch.int32, device=dev)
        vert_mesh = torch.empty(SV, dtype=torch.int32, device=dev)
        aux = torch.empty(2 * SV, dtype=torch.int32, device=dev)
        _lib.call("mrb_cubify_emit", B, Z, Y, X, _lib.ptr(ws), _lib.ptr(meta), SV, SF, E, _lib.ptr(verts),
                  _lib.ptr(faces), _lib.ptr(adj), _lib.ptr(rowptr), _lib.ptr(col32), _lib.ptr(vert_mesh),
                  _lib.ptr(aux))
    topo = MeshTopology(SV, E, rowptr, col32, symmetric=True, vert_mesh=vert_mesh)
    register(adj, topo)
    return verts, v_counts[:last], faces, f_counts[:last], adj, topo

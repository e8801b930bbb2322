"""Packed-mesh topology side-car.

The reference API passes adjacency around as a bare ``2 x E`` int64 COO tensor (``adj_index`` returned by
``Cubify.forward``, reference meshRCNN/layers.py:478,484, consumed by ``GraphConv.forward`` :47 and
``total_edge_length`` loss_functions.py:175).  The kernels want CSR with int32 columns.  ``Cubify`` already
produces that form on the device, so it registers it here keyed by the identity of the COO tensor it
returns; any other COO tensor is converted on the device the first time it is seen (no host sync).
"""
import weakref
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib


class MeshTopology:
    """CSR adjacency (row-sorted, int32) of a packed batch of meshes, plus the CSR of its transpose (the
    backward of the neighbour sum gathers along columns; for symmetric adjacency both are the same arrays)."""

    __slots__ = ("num_vertices", "num_edges", "rowptr", "col", "rowptr_t", "col_t", "symmetric", "vert_mesh",
                 "v_offsets", "row32", "__weakref__")

    def __init__(self, num_vertices, num_edges, rowptr, col, rowptr_t=None, col_t=None, symmetric=False,
                 vert_mesh=None, v_offsets=None, row32=None):
        self.num_vertices = int(num_vertices)
        self.num_edges = int(num_edges)
        self.rowptr = rowptr
        self.col = col
        self.symmetric = bool(symmetric)
        self.rowptr_t = rowptr if symmetric else rowptr_t
        self.col_t = col if symmetric else col_t
        self.vert_mesh = vert_mesh
        self.v_offsets = v_offsets
        self.row32 = row32


_REGISTRY: Dict[Tuple[int, int, int], Tuple[weakref.ref, MeshTopology]] = {}


def _key(adj: Tensor, n: int) -> Tuple[int, int, int]:
    return (adj.data_ptr(), adj.shape[1], n)


def register(adj: Tensor, topo: MeshTopology) -> None:
    if len(_REGISTRY) > 64:
        for k in [k for k, (r, _) in _REGISTRY.items() if r() is None]:
            del _REGISTRY[k]
    _REGISTRY[_key(adj, topo.num_vertices)] = (weakref.ref(adj), topo)


def lookup(adj: Tensor, num_vertices: int) -> Optional[MeshTopology]:
    hit = _REGISTRY.get(_key(adj, num_vertices))
    if hit is None:
        return None
    ref, topo = hit
    t = ref()
    if t is None or t is not adj and (t.data_ptr() != adj.data_ptr() or t._version != adj._version):
        return None
    return topo


def from_coo(adj: Tensor, num_vertices: int) -> MeshTopology:
    """CSR (+ transpose CSR) of an arbitrary ``2 x E`` int64 COO edge list, built on the device."""
    topo = lookup(adj, num_vertices)
    if topo is not None:
        return topo
    if not adj.is_cuda:
        raise RuntimeError("meshrcnn_b200: adjacency must be a CUDA tensor (no CPU fallback)")
    if adj.dim() != 2 or adj.shape[0] != 2:
        raise RuntimeError("meshrcnn_b200: adjacency must be 2 x E (COO)")
    adj_c = adj.contiguous().to(torch.int64)
    E = adj_c.shape[1]
    n = int(num_vertices)
    dev = adj.device
    out = []
    for transpose in (0, 1):
        rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        col = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
        ws = torch.empty(2 * (n + 1) + 2, dtype=torch.int32, device=dev)
        _lib.call("mrb_coo_to_csr", _lib.ptr(adj_c), E, n, transpose, _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(ws))
        out += [rowptr, col]
    topo = MeshTopology(n, E, out[0], out[1], out[2], out[3], symmetric=False)
    register(adj, topo)
    return topo

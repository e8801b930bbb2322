"""Mirror of reference ``utils/mesh_sampling.py`` (+ ``utils/process.normalize_mesh``) on the CUDA kernels."""
from typing import Optional

import torch
from torch import Tensor

from . import _lib, functional as F_


def surface_areas(vertex_positions: Tensor, mesh_faces: Tensor) -> Tensor:
    """Triangle areas |AB x AC| / 2 of one mesh -- reference utils/mesh_sampling.py:39-57."""
    return F_.face_areas(vertex_positions, mesh_faces, [vertex_positions.shape[0]], [mesh_faces.shape[0]])


def sample(vertex_positions: Tensor, mesh_faces: Tensor, num_points: float = 10e3, *, u: Optional[Tensor] = None,
           face_idx: Optional[Tensor] = None, xi2: Optional[Tensor] = None, xi1: Optional[Tensor] = None,
           seed: Optional[int] = None) -> Tensor:
    """Area-weighted surface samples of one mesh, centred and scaled into the unit ball -- reference
    utils/mesh_sampling.py:6-35.  Differentiable w.r.t. the vertex positions.  The keyword-only arguments inject
    the three random draws (parity tests); by default an in-kernel Philox stream seeded from torch's generator."""
    n = int(num_points)
    one = lambda t: None if t is None else t.reshape(1, n)
    cloud, _ = F_.sample_points(vertex_positions, mesh_faces, [vertex_positions.shape[0]], [mesh_faces.shape[0]], n,
                                u=one(u), face_idx=one(face_idx), xi2=one(xi2), xi1=one(xi1), seed=seed)
    return cloud[0]


def normalize_mesh(vertices: Tensor) -> Tensor:
    """Centre; if any |coordinate| > 1 divide by the largest vertex norm -- reference utils/process.py:7-20
    (forward only; inside ``sample`` the same kernel runs with its backward)."""
    F_._require_cuda(vertices, "normalize_mesh")
    v = F_._f32c(vertices.detach()).reshape(1, -1, 3)
    out = torch.empty_like(v)
    stats = torch.empty(1, 8, dtype=torch.float64, device=v.device)
    _lib.call("mrb_normalize_cloud_fwd", _lib.ptr(v), 1, v.shape[1], _lib.ptr(out), _lib.ptr(stats))
    return out[0]

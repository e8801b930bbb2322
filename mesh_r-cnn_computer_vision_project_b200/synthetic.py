"""Seeded synthetic inputs for the hot path.  Host-side torch only; shared by the
parity tests, ``bench.py`` and ``oracle/make_golden.py`` so that every implementation sees identical inputs.
Datasets / checkpoints are not available offline, so these stand in for ShapeNet / Pix3D batches.
"""
from typing import List, Sequence, Tuple

import torch
from torch import Tensor

SHAPENET_MAPS = ((256, 35, 35), (512, 18, 18), (1024, 9, 9), (2048, 5, 5))   # ResNet50 maps at 137x137 input
PIX3D_MAP = (256, 12, 12)                                                    # RoI features, pix3d_model.py


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def blob_voxels(B: int, V: int, seed: int = 0) -> Tensor:
    """Ellipsoid-sigmoid occupancy probabilities, B x V x V x V fp32: one connected blob per mesh; at
    threshold 0.2 about 18.5 % of the voxels are occupied."""
    g = _gen(seed)
    ax = torch.arange(V, dtype=torch.float32)
    zz, yy, xx = torch.meshgrid(ax, ax, ax, indexing="ij")
    grid = torch.stack([zz, yy, xx], dim=-1)
    out = torch.empty(B, V, V, V, dtype=torch.float32)
    for b in range(B):
        c = V / 2 + (torch.rand(3, generator=g) - 0.5) * 0.1 * V
        a = V * (0.22 + 0.12 * torch.rand(3, generator=g))
        d = (((grid - c) / a) ** 2).sum(-1).sqrt()
        out[b] = torch.sigmoid(6 * (1 - d))
    return out


def dense_voxels(B: int, V: int = 48, seed: int = 0) -> Tensor:
    """Uniform-random probabilities (Cubify stress, threshold 0.5 => density 0.5)."""
    return torch.rand(B, V, V, V, generator=_gen(seed))


def feature_maps(B: int, shapes: Sequence[Tuple[int, int, int]], seed: int = 0) -> List[Tensor]:
    g = _gen(seed + 17)
    return [torch.randn(B, c, h, w, generator=g) for (c, h, w) in shapes]


def in_frustum_positions(n: int, image_hw: int, seed: int = 0) -> Tensor:
    """Vertex positions whose projection (layers.py:557-558) lands strictly inside the image, so that the
    VertexAlign gather is actually exercised (pipeline coordinates all clamp to the border -- SURVEY finding 8)."""
    g = _gen(seed + 29)
    r = 0.44 if image_hw >= 224 else 0.09
    p2 = -(1 + 2 * torch.rand(n, generator=g))
    u1 = -0.44 + (r + 0.44) * torch.rand(n, generator=g)
    u0 = -0.44 + (r + 0.44) * torch.rand(n, generator=g)
    return torch.stack([-p2 * u0, p2 * u1, p2], dim=1).contiguous()


def sampling_randomness(B: int, n: int, seed: int) -> Tuple[Tensor, Tensor, Tensor]:
    """(u_face, xi2, xi1): three B x n uniform draws -- face selector (inverse CDF), and the two barycentric
    draws in the reference's order (mesh_sampling.py:20-21: xi2 first, then xi1 whose sqrt is used)."""
    g = _gen(seed + 43)
    return (torch.rand(B, n, generator=g), torch.rand(B, n, generator=g), torch.rand(B, n, generator=g))

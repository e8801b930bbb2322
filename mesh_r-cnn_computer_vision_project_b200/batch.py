"""Packed host->device batch container with the field layout of the reference ``data.dataloader.Batch``
(data/dataloader.py:11-77) -- the object ``batched_mesh_loss`` reads its ground truth from (``.meshes``,
``.vertice_index``, ``.face_index``) and the models read ``.images`` / ``.voxels`` / ``.backbone_targets`` from.

``meshes`` is a ``Mesh(vertices SV x 3, faces SF x 3)`` of all ground-truth meshes concatenated, faces holding per-mesh
LOCAL vertex ids; the three index lists split it back.  ``voxels`` are stacked and resampled to ``num_voxels``^3 like
``utils/process.py:24-41`` (adaptive max-pool down, nearest interpolation up)."""
from typing import List, Sequence, Union

import torch
from torch import Tensor
from torch.nn.functional import adaptive_max_pool3d, interpolate

from .serialization import Mesh


def resample_voxels(voxels: Tensor, N: int) -> Tensor:
    """B x V x V x V -> B x N x N x N (reference utils/process.py:24-41)."""
    assert voxels.ndim == 4, "expects batched input of shape BxVxVxV"
    M = voxels.shape[1]
    assert voxels.shape[1:] == torch.Size([M, M, M])
    if M > N:
        return adaptive_max_pool3d(voxels.to(torch.float32), N).to(voxels.dtype)
    if M < N:
        return interpolate(voxels.to(torch.float32).unsqueeze(1), size=N).squeeze(1).to(voxels.dtype)
    return voxels


class Batch:
    def __init__(self, images, voxels: Union[Tensor, Sequence[Tensor]], num_voxels: int, meshes: List[Mesh], backbone_targets):
        batched = voxels if isinstance(voxels, Tensor) else torch.stack(list(voxels))
        if batched.shape[1:] != torch.Size([num_voxels] * 3):
            batched = resample_voxels(batched, num_voxels)
        self.images = images
        self.voxels = batched
        self.meshes = Mesh(torch.cat([m.vertices for m in meshes]), torch.cat([m.faces for m in meshes]))
        self.mesh_index = [1 for _ in images]
        self.vertice_index = [m.vertices.shape[0] for m in meshes]
        self.face_index = [m.faces.shape[0] for m in meshes]
        self.backbone_targets = backbone_targets

    def to(self, *args, **kwargs) -> "Batch":
        if self.images is not None:
            if isinstance(self.images, (list, tuple)):
                self.images = type(self.images)([i.to(*args, **kwargs) for i in self.images])
            else:
                self.images = self.images.to(*args, **kwargs)
        if self.voxels is not None:
            self.voxels = self.voxels.to(*args, **kwargs)
        if self.meshes is not None:
            self.meshes = Mesh(self.meshes.vertices.to(*args, **kwargs), self.meshes.faces.to(*args, **kwargs))
        if self.backbone_targets is not None:
            self.backbone_targets = self.backbone_targets.to(*args, **kwargs)
        return self

    def pin_memory(self) -> "Batch":
        """Page-locked copies for asynchronous H2D transfers (the wire format of ``bench.py``'s end-to-end arm)."""
        pin = lambda t: t.pin_memory() if isinstance(t, Tensor) else t
        if isinstance(self.images, (list, tuple)):
            self.images = type(self.images)([pin(i) for i in self.images])
        else:
            self.images = pin(self.images)
        self.voxels = pin(self.voxels)
        self.meshes = Mesh(pin(self.meshes.vertices), pin(self.meshes.faces))
        self.backbone_targets = pin(self.backbone_targets)
        return self

    def __getitem__(self, idx) -> "Batch":
        if isinstance(idx, int):
            return self[idx:idx + 1]
        voxels = self.voxels[idx]
        meshes = [Mesh(v, f) for v, f in zip(self.meshes.vertices.split(self.vertice_index)[idx],
                                             self.meshes.faces.split(self.face_index)[idx])]
        return Batch(self.images[idx], voxels, voxels.shape[1], meshes, self.backbone_targets[idx])

    def __len__(self) -> int:
        return len(self.images)

"""Drop-in mirror of the reference ``meshRCNN/layers.py`` module API for the hot path: ``Cubify``,
``VertexAlign``, ``GraphConv``, ``ResGraphConv`` and the three refinement-stage classes -- same class names,
constructor arguments, ``forward`` argument orders, return tuples and ``state_dict`` keys, so that
``meshRCNN/shapenet_model.py:11-13,27-41,72-90`` and ``meshRCNN/pix3d_model.py:16-17,32-44,87-106`` work
unchanged after ``from meshrcnn_b200.layers import ...``.

Every numerical op runs in hand-written sm_100a CUDA behind the C ABI (``include/meshrcnn_b200.h``); there is
no CPU path: CPU tensors raise ``RuntimeError``.
"""
import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib, functional as F_, topology


class Cubify(nn.Module):
    """Voxel occupancy probabilities -> packed cuboid triangle meshes.

    Mirrors reference ``Cubify`` (meshRCNN/layers.py:342-484): ``forward(t)`` with ``t`` of shape B x Z x Y x X
    returns ``(vs, v_index, faces, f_index, adj_index)``: ``vs`` SV x 3 fp32 in voxel units rotated to
    (z, x, -y); ``v_index`` / ``f_index`` Python lists of per-mesh counts (truncated after the last non-empty
    mesh, like ``bincount``); ``faces`` SF x 3 int64 with per-mesh local vertex ids; ``adj_index`` 2 x E int64 with
    global ids sorted by (row, col).  Vertex / face / edge order equals the reference's (bit-exact).
    Raises ``ValueError("empty grid")`` when no voxel of the whole batch is occupied (layers.py:434-435).
    """

    def __init__(self, threshold: float = 0.5):
        super().__init__()
        self.threshold = threshold
        # state-dict compatibility with the reference ("cubify.kernel", "cubify.deltas", layers.py:365,401);
        # the CUDA path does not read them (the same tables live in csrc/cubify.cu as constants).
        nbr = [(-1, 0, 0), (1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, -1), (0, 0, 1)]
        kernel = torch.zeros(6, 1, 3, 3, 3)
        for d, (dz, dy, dx) in enumerate(nbr):
            kernel[d, 0, 1, 1, 1] = 1
            kernel[d, 0, 1 + dz, 1 + dy, 1 + dx] = -1
        corners = F_.CUBIFY_CORNERS
        deltas = torch.zeros(6, 4, 5)
        for d in range(6):
            for c in range(4):
                for a in range(3):
                    deltas[d, c, 2 + a] = corners[d][c][a] - 0.5
        self.register_buffer("kernel", kernel)
        self.register_buffer("deltas", deltas)

    def forward(self, t: Tensor):
        verts, v_index, faces, f_index, adj, _ = F_.cubify(t, float(self.threshold))
        return verts, v_index, faces, f_index, adj

"""Drop-in mirror of the mesh losses of reference ``meshRCNN/loss_functions.py`` on the CUDA kernels.

``batched_mesh_loss`` / ``mesh_loss`` keep the reference signatures (``batch`` is any object with ``.meshes``
(vertices, faces), ``.vertice_index`` and ``.face_index``, i.e. reference ``data.dataloader.Batch`` :21-36) and the
reference normalisation (sums over the batch divided by ``point_cloud_size`` only, :66,72).  Dense-matrix helpers
(``batched_point2point_distance``, ``batched_chamfer_distance(p2p)``, ``total_edge_length(p2p, adj)``) are kept as
thin compatibility functions for callers / tests that hold a distance matrix; ``mesh_loss`` never forms one.
"""
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import functional as F_


def voxel_loss(voxel_prediction: Tensor, voxel_gts: Tensor) -> Tensor:
    """mean BCE on the voxel occupancy probabilities -- reference loss_functions.py:10-14, one fused pass with fp64
    accumulation and torch's clamp of the log terms (csrc/voxel.cu)."""
    return F_.voxel_bce(voxel_prediction, voxel_gts, from_logits=False)[0]


def voxel_loss_with_logits(voxel_logits: Tensor, voxel_gts: Tensor, return_probs: bool = False):
    """The same loss taken from the voxel head's LOGITS: the ``nn.Sigmoid`` that ends the reference's
    ``VoxelBranch`` (layers.py:505) is evaluated inside the loss kernel, so training never materialises the probability grid
    (``Cubify(threshold)(logits, from_logits=True)`` thresholds the same sigmoid in its first kernel).  Returns the loss, or
    ``(loss, probabilities)`` with ``return_probs=True`` (eval output dict, shapenet_model.py:66)."""
    loss, probs = F_.voxel_bce(voxel_logits, voxel_gts, from_logits=True, want_probs=return_probs)
    return (loss, probs) if return_probs else loss


def batched_mesh_loss(vertex_positions_pred: List[Tensor], mesh_faces_pred: Tensor, pred_adjacency: Tensor,
                      vertices_per_sample_pred: List[int], faces_per_sample_pred: List[int], batch,
                      point_cloud_size: float = 10e3, num_neighbours_for_normal_loss: int = 10,
                      randomness=None, gt_clouds=None) -> Tuple[Tensor, Tensor, Tensor]:
    """Sum over the refinement stages of (chamfer, normal, edge) -- reference loss_functions.py:17-35.
    ``randomness``: optional list (one entry per stage) of ``(rnd_pred, rnd_gt)`` injected draws, see ``mesh_loss``;
    ``gt_clouds``: optional list of already drawn ground-truth samples, one per stage."""
    terms = [mesh_loss(pos, mesh_faces_pred, pred_adjacency, vertices_per_sample_pred, faces_per_sample_pred, batch,
                       point_cloud_size, num_neighbours_for_normal_loss,
                       randomness=None if randomness is None else randomness[s],
                       gt_cloud=None if gt_clouds is None else gt_clouds[s])
             for s, pos in enumerate(vertex_positions_pred)]
    return tuple(F_.weighted_scalar_sum([t[i] for t in terms]) for i in range(3))     # one launch per term (was 2 adds each)


def mesh_loss(vertex_positions_pred: Tensor, mesh_faces_pred: Tensor, pred_adjacency: Tensor,
              vertices_per_sample_pred: List[int], faces_per_sample_pred: List[int], batch,
              point_cloud_size: float = 10e3, num_neighbours_for_normal_loss: int = 10,
              randomness=None, gt_cloud: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """(chamfer, normal, edge) of one stage -- reference loss_functions.py:40-74.
    ``gt_cloud``: the stage's ground-truth sample if the caller already drew it (``sample_gt_cloud``; the refinement head
    samples the static GT meshes ahead of the predicted mesh, with the seeds in the order of the lazy path).

    ``randomness = (rnd_pred, rnd_gt)`` with ``rnd_* = dict(u=|face_idx=, xi2=, xi1=)`` (B x n tensors) injects the
    sampling draws of the predicted / ground-truth clouds; default: fresh in-kernel Philox draws for both, the GT
    cloud being re-sampled at every call like the reference (:57-59)."""
    n = int(point_cloud_size)
    k = int(num_neighbours_for_normal_loss)
    rnd_pred, rnd_gt = randomness if randomness is not None else ({}, {})

    edge_loss = F_.edge_length(vertex_positions_pred, pred_adjacency)                               # :47-48

    cloud_pred, _ = F_.sample_points(vertex_positions_pred, mesh_faces_pred, vertices_per_sample_pred,
                                     faces_per_sample_pred, n, **rnd_pred)                          # :51-53
    # :57-59 -- the GT cloud is re-sampled at every call like the reference; the area CDF of the static GT meshes is
    # computed once per batch object (device-resident cache, functional.cached_face_cdf)
    if gt_cloud is None:
        gt_cloud = sample_gt_cloud(batch, n, **rnd_gt)
    cloud_gt = gt_cloud

    # :62-66,141  (loss_p + loss_gt) / point_cloud_size as one scalar; :69-72  -(nd_p + nd_gt) / point_cloud_size likewise
    chamfer_loss, idx_p, idx_gt, knn_p, knn_gt = F_.chamfer_total(cloud_pred, cloud_gt, k, 1.0 / point_cloud_size)
    normal_loss = F_.normal_total(cloud_pred, cloud_gt, knn_p, knn_gt, idx_p, idx_gt, -1.0 / point_cloud_size)
    return chamfer_loss, normal_loss, edge_loss


def sample_gt_cloud(batch, point_cloud_size: float = 10e3, **randomness) -> Tensor:
    """B x n x 3 normalised surface sample of the ground-truth meshes of ``batch`` (reference loss_functions.py:57-59)."""
    pos_gt, faces_gt = batch.meshes
    cloud, _ = F_.sample_points(pos_gt, faces_gt, batch.vertice_index, batch.face_index, int(point_cloud_size), cdf_owner=batch,
                                **randomness)
    return cloud


def batched_mesh_sampling(vertex_positions: Tensor, mesh_faces: Tensor, vertices_per_sample: List[int],
                          faces_per_sample: List[int], num_points: float = 10e3, **randomness) -> Tensor:
    """B x n x 3 normalised surface samples of a packed batch -- reference loss_functions.py:80-89."""
    cloud, _ = F_.sample_points(vertex_positions, mesh_faces, vertices_per_sample, faces_per_sample, int(num_points),
                                **randomness)
    return cloud


def batched_point2point_distance(pt0: Tensor, pt1: Optional[Tensor] = None) -> Tensor:
    """Dense |a_i - b_j|^2 matrix -- reference loss_functions.py:192-220.  Compatibility helper for callers that
    want the matrix itself (tests, small inputs); the loss path never calls it."""
    if pt0.ndim == 2:
        pt0 = pt0.unsqueeze(0)
    if pt1 is None:
        pt1 = pt0
    elif pt1.ndim == 2:
        pt1 = pt1.unsqueeze(0)
    diff = pt0.unsqueeze(2) - pt1.unsqueeze(1)
    return (diff * diff).sum(-1)


def batched_chamfer_distance(p2p_distance: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """(loss_1, idx1, loss_2, idx2) from a dense distance matrix -- reference loss_functions.py:93-102
    (compatibility helper; see ``chamfer_distance`` for the fused path)."""
    mins, idx1 = torch.min(p2p_distance, 2)
    loss_1 = torch.sum(mins)
    mins, idx2 = torch.min(p2p_distance, 1)
    return loss_1, idx1, torch.sum(mins), idx2


def chamfer_distance(p: Tensor, q: Tensor, k: int = 0):
    """Fused replacement of ``batched_chamfer_distance(batched_point2point_distance(p, q))``: returns
    (loss_1, idx1, loss_2, idx2[, knn_p, knn_q]) without forming the B x P x Q matrix."""
    l1, l2, i1, i2, kp, kq = F_.chamfer_knn(p, q, k)
    if k:
        return l1, i1.long(), l2, i2.long(), kp, kq
    return l1, i1.long(), l2, i2.long()


def batched_normal_distance(p: Tensor, pgt: Tensor, p2p_distance: Optional[Tensor], idx_p: Tensor, idx_gt: Tensor,
                            k: int = 4) -> Tuple[Tensor, Tensor]:
    """Normal-consistency sums -- reference loss_functions.py:107-126.  ``p2p_distance`` is accepted for signature
    compatibility and ignored: the k-NN sets are recomputed by the fused search."""
    _, _, _, _, knn_p, knn_gt = F_.chamfer_knn(p.detach(), pgt.detach(), k)
    return F_.normal_distance(p, pgt, knn_p, knn_gt, idx_p, idx_gt)


def compute_normals(pt: Tensor, p2p_distance: Optional[Tensor] = None, k: int = 10, other: Optional[Tensor] = None,
                    knn: Optional[Tensor] = None) -> Tensor:
    """Per-point PCA 'normals' -- reference loss_functions.py:129-170.  Pass the other cloud (``other``) or the
    k-NN index sets (``knn``); a dense ``p2p_distance`` (rows = pt, columns = other cloud) is also accepted."""
    if knn is None:
        if other is not None:
            knn = F_.chamfer_knn(pt.detach(), other.detach(), k)[4]
        elif p2p_distance is not None:
            knn = p2p_distance.topk(k, dim=2, largest=False, sorted=False).indices
        else:
            raise RuntimeError("compute_normals needs `other`, `knn` or a distance matrix")
    return F_.compute_normals(pt, knn)


def total_edge_length(p2p_distance: Tensor, vertex_adjacency: Tensor) -> Tensor:
    """Mean squared edge length from a dense distance matrix -- reference loss_functions.py:175-189
    (compatibility helper; ``edge_length(pos, adj)`` is the O(E) kernel the loss uses)."""
    masked = p2p_distance[vertex_adjacency[0], vertex_adjacency[1]]
    return masked.sum() / masked.shape[0]


def edge_length(vertex_positions: Tensor, vertex_adjacency: Tensor) -> Tensor:
    return F_.edge_length(vertex_positions, vertex_adjacency)

"""Eval-side helpers next to the hot path: the metrics of the reference ``utils/metrics.py`` and the
detection-selection + adjacency rebuild of ``utils/eval_utils.py:12-90``.  Device-agnostic tensor code (works on the
GPU tensors the refinement head returns in eval mode and on CPU tensors); no custom kernels are involved: the rebuilt
adjacency is handed to the CSR path through ``topology.from_coo`` like any foreign COO tensor.

Reference semantics kept: ``f_score`` works on a confusion matrix c[i, j] = #(predicted i, ground truth j) and returns
percentages with the 1e-8 guards of :20-26; ``mesh_precision_recall`` zeroes ALL true positives through the scalar-mask
assignment ``tp[f1 <= 0.5] = 0`` when the running F0.3 is <= 0.5 (:53-60) and integrates precision over recall with
the trapezoid rule (sklearn ``auc``: the recall values must be monotonic)."""
from typing import List, Sequence, Tuple

import torch
from torch import Tensor


def f_score(confusion: Tensor, beta: float = 1.0) -> Tensor:
    tp = confusion.diagonal()
    precision = 100 * (tp / (1e-8 + confusion.sum(1)))
    recall = 100 * (tp / (1e-8 + confusion.sum(0)))
    return (1 + beta ** 2) * precision * recall / (1e-8 + recall + (beta ** 2) * precision)


def _auc(x: Tensor, y: Tensor) -> float:
    """Trapezoid area under y(x) for monotonic x (what ``sklearn.metrics.auc`` computes; decreasing x flips the sign)."""
    if x.numel() < 2:
        raise ValueError("At least 2 points are needed to compute area under curve, but x.shape = %d" % x.numel())
    dx = x[1:] - x[:-1]
    if bool((dx < 0).any()):
        if bool((dx <= 0).all()):
            sign = -1.0
        else:
            raise ValueError("x is neither increasing nor decreasing : {}.".format(x.tolist()))
    else:
        sign = 1.0
    return sign * float((dx.double() * (y[1:] + y[:-1]).double() / 2).sum())


def mesh_precision_recall(confusion: Tensor, f1_score: float) -> float:
    tp = confusion.diagonal().clone()
    if f1_score <= 0.5:                     # reference: tp[f1_score <= 0.5] = 0 with a Python scalar -> all or nothing
        tp = torch.zeros_like(tp)
    precision = 100 * (tp / (1e-8 + confusion.sum(1)))
    recall = 100 * (tp / (1e-8 + confusion.sum(0)))
    return _auc(recall, precision)


def box_iou(a: Tensor, b: Tensor) -> Tensor:
    """Pairwise IoU of (x1, y1, x2, y2) boxes, N x 4 and M x 4 -> N x M."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area_a[:, None] + area_b[None, :] - inter)


def calc_precision_box(boxes: Sequence[Tensor], gt_boxes: Sequence[Tensor]) -> float:
    hits = sum(1 for g, p in zip(gt_boxes, boxes) if box_iou(g, p.unsqueeze(0))[0][0] > 0.5)
    return hits / len(boxes)


def calc_precision_mask(masks: Sequence[Tensor], gt_masks: Sequence[Tensor]) -> float:
    hits = 0
    for mask, gt in zip(masks, gt_masks):
        m, g = (mask > 0.5).to(torch.int32), gt.to(torch.int32)
        if torch.sum(m & g) / torch.sum(m | g) > 0.5:
            hits += 1
    return hits / len(masks)


def get_max_box(boxes: Tensor, gt_box: Tensor) -> Tuple[Tensor, Tensor]:
    idx = torch.argmax(box_iou(boxes, gt_box), dim=0)[0]
    return boxes[idx], idx


def faces_to_adjacency(faces_global: Tensor) -> Tensor:
    """2 x E symmetric COO adjacency (sorted by (row, col), duplicates removed) from SF x 3 faces with GLOBAL vertex ids:
    the three edges of every triangle in both directions (reference eval_utils.py:77-88 == layers.py:469-478)."""
    ft = faces_global.t()
    i, j = torch.cat([ft[:2], ft[1:], ft[::2]], dim=1)
    i, j = torch.cat([i, j], dim=0), torch.cat([j, i], dim=0)
    return torch.stack([i, j], dim=0).unique(dim=1)


def get_only_max(max_indexes: Sequence[int], voxels: Tensor, vertex_positions: List[Tensor], faces: Tensor,
                 vertice_index: List[int], face_index: List[int], mesh_index: List[int]):
    """Keeps, per image, the detection ``max_indexes[img]`` of its ``mesh_index[img]`` meshes and rebuilds the packed batch:
    (voxels, per-stage positions, faces (local ids), adjacency (global ids), vertice_index, face_index)."""
    vxls = torch.stack([g[idx] for g, idx in zip(voxels.split(mesh_index), max_indexes)])
    fs = faces.split(face_index)
    res_fs, i = [], 0
    for n, idx in zip(mesh_index, max_indexes):
        res_fs.append(fs[i:i + n][idx])
        i += n
    res_f_index = [f.size(0) for f in res_fs]
    res_vs, res_v_index = [], []
    for j, stage in enumerate(vertex_positions):
        vs, picked, i = stage.split(vertice_index), [], 0
        for n, idx in zip(mesh_index, max_indexes):
            picked.append(vs[i:i + n][idx])
            i += n
        if j == 0:
            res_v_index = [v.size(0) for v in picked]
        res_vs.append(torch.cat(picked))
    offsets, run = [], 0
    for n in res_v_index:
        offsets.append(run)
        run += n
    adj = faces_to_adjacency(torch.cat([f + off for f, off in zip(res_fs, offsets)]))
    return vxls, res_vs, torch.cat(res_fs), adj, res_v_index, res_f_index

"""TEST INFRASTRUCTURE ONLY -- fixtures for the eval-side helpers (``tests/golden/eval_expected.npz``) from the UNMODIFIED
reference: ``utils/metrics.py`` is loaded by file path; ``get_only_max`` is taken out of ``utils/eval_utils.py`` by its AST
node at run time (the module itself imports the training stack), executed, never copied.

    python -m oracle.make_golden_eval       # dev container only (needs /root/reference)
"""
import ast
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402


def inputs():
    g = torch.Generator().manual_seed(7)
    conf = torch.randint(0, 9, (5, 5), generator=g).float() + 3 * torch.eye(5)
    mesh_index = [2, 1, 3]
    v_index = [4, 6, 5, 3, 7, 4]
    f_index = [3, 5, 4, 2, 6, 3]
    faces = torch.cat([torch.randint(0, n, (f, 3), generator=g) for n, f in zip(v_index, f_index)])
    positions = [torch.rand(sum(v_index), 3, generator=g) for _ in range(3)]
    voxels = torch.rand(6, 4, 4, 4, generator=g)
    max_idx = [1, 0, 2]
    boxes = torch.tensor([[0., 0., 10., 10.], [2., 2., 8., 9.], [20., 20., 30., 31.]])
    gt_box = torch.tensor([[1., 1., 9., 9.]])
    masks = [torch.rand(6, 6, generator=g) for _ in range(4)]
    gt_masks = [(torch.rand(6, 6, generator=g) > 0.4) for _ in range(4)]
    return conf, mesh_index, v_index, f_index, faces, positions, voxels, max_idx, boxes, gt_box, masks, gt_masks


def main():
    spec = importlib.util.spec_from_file_location("ref_metrics", os.path.join(ref_import.REF_ROOT, "utils", "metrics.py"))
    met = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(met)
    src = open(os.path.join(ref_import.REF_ROOT, "utils", "eval_utils.py")).read()
    ns = {"torch": torch, "np": np, "box_iou": met.box_iou}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in ("get_only_max", "get_max_box"):
            exec(compile(ast.Module([node], []), "eval_utils.py", "exec"), ns)
    conf, mesh_index, v_index, f_index, faces, positions, voxels, max_idx, boxes, gt_box, masks, gt_masks = inputs()
    out = {"f_%d" % f: met.f_score(conf, f / 10).numpy() for f in (1, 3, 5, 10)}
    # sklearn's auc needs monotonic recall values: a confusion matrix with increasing per-class recall, and the random one
    # (for which the reference raises ValueError)
    mono = torch.diag(torch.arange(1., 6.)) + torch.roll(torch.diag(10 - torch.arange(1., 6.)), 1, 0)
    out["conf_mono"] = mono.numpy()
    out["ap_hi"] = np.float64(met.mesh_precision_recall(mono.clone(), 0.9))
    out["ap_lo"] = np.float64(met.mesh_precision_recall(mono.clone(), 0.2))
    try:
        met.mesh_precision_recall(conf.clone(), 0.9)
        out["random_conf_raises"] = np.int64(0)
    except ValueError:
        out["random_conf_raises"] = np.int64(1)
    vx, vs, fs, adj, vi, fi = ns["get_only_max"](max_idx, voxels, positions, faces, v_index, f_index, mesh_index)
    out.update(vx=vx.numpy(), fs=fs.numpy(), adj=adj.numpy(), vi=np.array(vi), fi=np.array(fi))
    for s, v in enumerate(vs):
        out["vs%d" % s] = v.numpy()
    mb, mi = ns["get_max_box"](boxes, gt_box)
    out.update(max_box=mb.numpy(), max_box_idx=np.int64(mi), p_box=np.float64(met.calc_precision_box([boxes[0], boxes[2]], [gt_box, gt_box])),
               p_mask=np.float64(met.calc_precision_mask(masks, gt_masks)))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "eval_expected.npz"), **out)
    print("wrote eval_expected.npz", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()

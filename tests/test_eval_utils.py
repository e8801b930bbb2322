"""Eval-side helpers against outputs of the UNMODIFIED reference (``oracle/make_golden_eval.py``): metrics of
``utils/metrics.py`` and the detection selection + adjacency rebuild of ``utils/eval_utils.py:12-90``.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle.make_golden_eval import inputs

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval_expected.npz"))


def test_f_scores_and_mesh_ap():
    from meshrcnn_b200 import eval_utils as E
    conf = inputs()[0]
    for f in (1, 3, 5, 10):
        assert np.allclose(E.f_score(conf, f / 10).numpy(), G["f_%d" % f], rtol=1e-6, atol=0)
    mono = torch.from_numpy(G["conf_mono"])
    assert abs(E.mesh_precision_recall(mono, 0.9) - float(G["ap_hi"])) <= 1e-6 * abs(float(G["ap_hi"]))
    assert E.mesh_precision_recall(mono, 0.2) == float(G["ap_lo"]) == 0.0          # all true positives zeroed (:56)
    assert int(G["random_conf_raises"]) == 1
    with pytest.raises(ValueError):                                               # like sklearn.metrics.auc in the reference
        E.mesh_precision_recall(conf, 0.9)


def test_detection_selection_and_adjacency_rebuild():
    from meshrcnn_b200 import eval_utils as E
    conf, mesh_index, v_index, f_index, faces, positions, voxels, max_idx, boxes, gt_box, masks, gt_masks = inputs()
    vx, vs, fs, adj, vi, fi = E.get_only_max(max_idx, voxels, positions, faces, v_index, f_index, mesh_index)
    assert np.array_equal(vx.numpy(), G["vx"]) and np.array_equal(fs.numpy(), G["fs"]) and np.array_equal(adj.numpy(), G["adj"])
    assert vi == G["vi"].tolist() and fi == G["fi"].tolist()
    for s, v in enumerate(vs):
        assert np.array_equal(v.numpy(), G["vs%d" % s])
    # the rebuilt adjacency is symmetric, sorted by (row, col) and duplicate free -- what topology.from_coo expects
    key = adj[0] * (int(adj.max()) + 1) + adj[1]
    assert bool((key[1:] > key[:-1]).all())
    assert torch.equal(E.faces_to_adjacency(torch.tensor([[0, 1, 2]])), torch.tensor([[0, 0, 1, 1, 2, 2], [1, 2, 0, 2, 0, 1]]))
    mb, mi = E.get_max_box(boxes, gt_box)
    assert np.array_equal(mb.numpy(), G["max_box"]) and int(mi) == int(G["max_box_idx"])
    assert E.calc_precision_box([boxes[0], boxes[2]], [gt_box, gt_box]) == float(G["p_box"])
    assert E.calc_precision_mask(masks, gt_masks) == float(G["p_mask"])

"""CPU: the oracle restatement (oracle/cubify_np.py, oracle/mesh_ops.py) against the committed outputs of the
UNMODIFIED reference (tests/golden/*.npz, produced by oracle/make_golden.py in the dev container), including the
reference's own shipped golden pair shapenet_ex/00_voxel_obj0.npy -> 00_mesh_stage0_obj_0.obj."""
import numpy as np
import pytest
import torch

from oracle import cubify_np, mesh_ops

CASES = ["rand10", "ragged", "empty_mid_tail", "single_voxel", "solid3", "at_threshold", "blob16", "dense24"]


def T(x, dt=torch.float64):
    t = torch.as_tensor(np.asarray(x))
    return t.to(dt) if t.is_floating_point() else t


def test_cubify_oracle_reproduces_shipped_golden(golden):
    g = golden("cubify_shapenet_ex")
    shape = tuple(g["voxel_shape"])
    vox = np.unpackbits(g["voxel_bits"])[:np.prod(shape)].reshape(shape).astype(np.float32)
    assert int(vox.sum()) == 1912
    v, vi, f, fi, adj = cubify_np.cubify(vox[None], 0.5)
    assert v.shape == (2629, 3) and f.shape == (4896, 3)
    assert np.array_equal(v, g["verts"]) and np.array_equal(f, g["faces"]) and np.array_equal(adj, g["adj"])


@pytest.mark.parametrize("name", CASES)
def test_cubify_oracle_vs_reference_outputs(golden, name):
    g = golden("cubify_cases")
    v, vi, f, fi, adj = cubify_np.cubify(g[name + "__in"], float(g[name + "__th"]))
    assert np.array_equal(v, g[name + "__verts"]) and vi == g[name + "__v_index"].tolist()
    assert np.array_equal(f, g[name + "__faces"]) and fi == g[name + "__f_index"].tolist()
    assert np.array_equal(adj, g[name + "__adj"])


def test_cubify_oracle_edge_cases():
    with pytest.raises(ValueError, match="empty grid"):
        cubify_np.cubify(np.zeros((2, 4, 4, 4), np.float32), 0.5)
    with pytest.raises(ValueError, match="empty grid"):
        cubify_np.cubify(np.full((1, 3, 3, 3), 0.5, np.float32), 0.5)       # p == th is empty (strict >)
    t = np.zeros((3, 4, 4, 4), np.float32)
    t[0, 1, 1, 1] = t[2, 2, 2, 2] = 1
    v, vi, f, fi, adj = cubify_np.cubify(t, 0.5)
    assert vi == [8, 0, 8] and fi == [12, 0, 12] and adj.shape[1] == 92      # single voxel: 8 v, 12 f, 46 directed edges
    t[2] = 0
    assert cubify_np.cubify(t, 0.5)[1] == [8]                                # trailing empties truncate the lists


@pytest.mark.parametrize("name,nmaps", [("pix", 1), ("shp", 4), ("randint", 2), ("border", 1)])
def test_vert_align_oracle(golden, name, nmaps):
    g = golden("vert_align")
    hw = int(g[name + "__hw"])
    fm = [T(g["%s__fm%d" % (name, i)], torch.float32).requires_grad_() for i in range(nmaps)]
    pos = T(g[name + "__pos"], torch.float32)
    out = mesh_ops.vert_align(fm, pos, g[name + "__vpm"].tolist(), [(hw, hw)] * 3, [1, 1, 1])
    assert np.array_equal(out.detach().numpy(), g[name + "_f32__out"])
    grads = torch.autograd.grad((out * T(g[name + "_f32__gout"], torch.float32)).sum(), fm)
    for i, gr in enumerate(grads):
        assert torch.allclose(gr, T(g["%s_f32__gfm%d" % (name, i)], torch.float32), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["gc_19_16", "gc_16_3", "gc_35_24"])
def test_graphconv_oracle(golden, name):
    g = golden("graphconv_stages")
    x = T(g[name + "__x"]).requires_grad_()
    w0, w1 = T(g[name + "__w0"]).requires_grad_(), T(g[name + "__w1"]).requires_grad_()
    out = mesh_ops.graph_conv(x, T(g["adj"]).long(), w0, w1)
    assert torch.allclose(out, T(g[name + "_f64__out"]), rtol=1e-12, atol=1e-12)
    gx, gw0, gw1 = torch.autograd.grad((out * T(g[name + "_f64__gout"])).sum(), [x, w0, w1])
    assert torch.allclose(gx, T(g[name + "_f64__gx"]), rtol=1e-10, atol=1e-12)
    assert torch.allclose(gw0, T(g[name + "_f64__gw0"]), rtol=1e-10, atol=1e-12)
    assert torch.allclose(gw1, T(g[name + "_f64__gw1"]), rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("cls", ["ResVertixRefineShapenet", "VertixRefineShapeNet", "VertixRefinePix3D"])
@pytest.mark.parametrize("use_feat", [0, 1])
def test_stage_oracle(golden, cls, use_feat):
    g = golden("graphconv_stages")
    tag = "%s_%d" % (cls, use_feat)
    sd = {k[len(tag) + 6:]: T(v) for k, v in g.items() if k.startswith(tag + "__sd__")}
    nmaps = 1 if cls == "VertixRefinePix3D" else 4
    fm = [T(g["%s__fm%d" % (cls, i)]) for i in range(nmaps)]
    # the fp64 reference run used the unrounded fp64 maps; the fixture stores them as fp32 -> compare at fp32 accuracy
    hw = int(g[cls + "__hw"])
    feats = T(g[cls + "__feats"]) if use_feat else None
    new_pos, new_feat = mesh_ops.STAGES[cls](sd, g["v_index"].tolist(), fm[0] if nmaps == 1 else fm, T(g["adj"]).long(),
                                             T(g[cls + "__pos"]), [(hw, hw)] * 2, feats=feats)
    assert torch.allclose(new_pos, T(g[tag + "_f64__new_pos"]), rtol=1e-5, atol=1e-5)
    assert torch.allclose(new_feat, T(g[tag + "_f64__new_feat"]), rtol=1e-5, atol=1e-5)


def test_losses_oracle(golden):
    g = golden("sampling_losses")
    v_index, f_index = g["v_index"].tolist(), g["f_index"].tolist()
    pos = T(g["pos"]).requires_grad_()
    faces, adj = T(g["faces"]).long(), T(g["adj"]).long()
    rp = (T(g["fi_p"]).long(), T(g["xi2_p"]), T(g["xi1_p"]))
    rg = (T(g["fi_g"]).long(), T(g["xi2_g"]), T(g["xi1_g"]))
    ch, nl, ed, inter = mesh_ops.mesh_loss_with(pos, faces, adj, v_index, f_index, T(g["gt_pos"]), T(g["gt_faces"]).long(),
                                                g["gt_v_index"].tolist(), g["gt_f_index"].tolist(), rp, rg,
                                                float(g["n_points"]), int(g["k"]))
    assert torch.allclose(ch, T(g["f64__chamfer"]), rtol=1e-10)
    assert torch.allclose(ed, T(g["f64__edge"]), rtol=1e-10)
    assert torch.allclose(nl, T(g["f64__normal"]), rtol=1e-6)        # LAPACK eigenvector signs: same library, same signs
    assert torch.allclose(inter["cloud_pred"], T(g["f64__cloud"]), rtol=1e-10, atol=1e-12)
    assert np.array_equal(inter["idx_p"].numpy(), g["f64__idx_p"])
    gch, = torch.autograd.grad(ch, pos, retain_graph=True)
    assert torch.allclose(gch, T(g["f64__chamfer_gpos"]), rtol=1e-8, atol=1e-12)
    ged, = torch.autograd.grad(ed, pos, retain_graph=True)
    assert torch.allclose(ged, T(g["f64__edge_gpos"]), rtol=1e-8, atol=1e-12)
    # inverse-CDF face draw used for the injected face indices
    fi = torch.stack([mesh_ops.face_cdf_draw(v, f, T(g["u_p"])[b]) for b, (v, f) in
                      enumerate(zip(pos.detach().split(v_index), faces.split(f_index)))])
    assert np.array_equal(fi.numpy(), g["fi_p"])


def test_known_answers_of_the_reference_tests():
    """tests/test_layers.py:16-26,58-74 and tests/test_loss_functions.py:14-55,76-96,100-125 restated on the oracle."""
    a = torch.tensor([[1., 2, 3], [4, 5, 6], [7, 8, 9]])
    ei = torch.tensor([[0, 0, 1, 2], [1, 2, 1, 0]])
    assert torch.equal(mesh_ops.aggregate_neighbours(ei, a), torch.tensor([[11., 13, 15], [4, 5, 6], [1, 2, 3]]))
    adj = torch.tensor([[0, 1, 0], [1, 0, 1], [0, 1, 0]]).nonzero().t()
    out = mesh_ops.graph_conv(torch.arange(9.).reshape(3, 3), adj, torch.ones(3, 6), torch.ones(3, 6))
    assert torch.equal(out, torch.tensor([15., 36, 33]).view(3, 1).expand(3, 6))
    x = torch.arange(15.).reshape(5, 3)
    d = mesh_ops.p2p_distance(x).squeeze()
    assert d[0, 4] == 432 and d[1, 3] == 108 and torch.equal(d, d.t())
    pt0, pt1 = torch.arange(30.).reshape(1, 10, 3), torch.arange(21.).reshape(1, 7, 3) + 1
    l0, i0, l1, i1 = mesh_ops.chamfer(mesh_ops.p2p_distance(pt0, pt1))
    assert l0.item() == 300 and l1.item() == 21
    v = torch.tensor([[0, 0, 0], [1, 0, 0], [1, 1, 1], [0, 0, 2], [0, 2, 0], [0, 1, 5], [2, 2, 2], [2, 7, 0], [2, 3, 5],
                      [2, 7, 8], [0, 3, 2]], dtype=torch.float32)
    f = torch.tensor([[1, 2, 8], [3, 4, 5], [0, 1, 7], [6, 9, 10]])
    assert torch.allclose(mesh_ops.surface_areas(v, f), torch.tensor([1.22474, 4.0, 3.5, 8.3666]), rtol=1e-5)

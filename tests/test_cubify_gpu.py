"""Cubify CUDA path vs the oracle / the reference's goldens -- bit-exact, values AND order."""
import numpy as np
import pytest
import torch

from oracle import cubify_np

pytestmark = pytest.mark.gpu

CASES = ["rand10", "ragged", "empty_mid_tail", "single_voxel", "solid3", "at_threshold", "blob16", "dense24"]


def _run(t, th):
    from meshrcnn_b200.layers import Cubify
    mod = Cubify(th).cuda()
    vs, vi, f, fi, adj = mod(torch.from_numpy(np.ascontiguousarray(t)).cuda())
    return vs.cpu().numpy(), vi, f.cpu().numpy(), fi, adj.cpu().numpy()


def _assert_same(got, want):
    assert got[1] == list(want[1]) and got[3] == list(want[3])
    assert got[0].dtype == np.float32 and np.array_equal(got[0], want[0])
    assert got[2].dtype == np.int64 and np.array_equal(got[2], want[2])
    assert got[4].dtype == np.int64 and np.array_equal(got[4], want[4])


def test_shapenet_ex_golden(lib, golden):
    """The reference's own shipped pair shapenet_ex/00_voxel_obj0.npy -> 00_mesh_stage0_obj_0.obj."""
    g = golden("cubify_shapenet_ex")
    shape = tuple(g["voxel_shape"])
    vox = np.unpackbits(g["voxel_bits"])[:np.prod(shape)].reshape(shape).astype(np.float32)
    vs, vi, f, fi, adj = _run(vox[None], 0.5)
    assert np.array_equal(vs, g["verts"]) and np.array_equal(f, g["faces"])
    assert np.array_equal(adj, g["adj"].astype(np.int64))
    assert vi == [g["verts"].shape[0]] and fi == [g["faces"].shape[0]]


@pytest.mark.parametrize("name", CASES)
def test_reference_cases(lib, golden, name):
    g = golden("cubify_cases")
    got = _run(g[name + "__in"], float(g[name + "__th"]))
    want = (g[name + "__verts"], g[name + "__v_index"].tolist(), g[name + "__faces"].astype(np.int64),
            g[name + "__f_index"].tolist(), g[name + "__adj"].astype(np.int64))
    _assert_same(got, want)


def test_empty_grid_raises(lib):
    from meshrcnn_b200.layers import Cubify
    with pytest.raises(ValueError, match="empty grid"):
        Cubify(0.5)(torch.zeros(2, 4, 4, 4, device="cuda"))
    with pytest.raises(ValueError, match="empty grid"):
        Cubify(0.5)(torch.full((1, 4, 4, 4), 0.5, device="cuda"))     # p == threshold is empty (strict >)


def test_cpu_tensor_rejected(lib):
    from meshrcnn_b200.layers import Cubify
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Cubify(0.5)(torch.ones(1, 4, 4, 4))


@pytest.mark.parametrize("B,V,kind,th", [(4, 32, "blob", 0.2), (32, 24, "blob", 0.2), (3, 48, "blob", 0.2),
                                         (2, 48, "dense", 0.5), (5, 13, "dense", 0.8)])
def test_vs_oracle_at_config_sizes(lib, B, V, kind, th):
    from meshrcnn_b200 import synthetic
    t = (synthetic.blob_voxels(B, V, 0) if kind == "blob" else synthetic.dense_voxels(B, V, 0)).numpy()
    _assert_same(_run(t, th), cubify_np.cubify(t, th))


def test_stress_config4_properties(lib):
    """BASELINE config 4 (64 dense 48^3 grids): checked through size-independent invariants + a 2-mesh oracle slice."""
    from meshrcnn_b200 import synthetic
    from meshrcnn_b200.layers import Cubify
    t = synthetic.dense_voxels(64, 48, 0)
    vs, vi, f, fi, adj = Cubify(0.5)(t.cuda())
    assert len(vi) == 64 and len(fi) == 64 and sum(vi) == vs.shape[0] and sum(fi) == f.shape[0]
    # adjacency strictly increasing in (row, col), symmetric, no self loops
    key = adj[0] * vs.shape[0] + adj[1]
    assert bool((key[1:] > key[:-1]).all())
    assert bool((adj[0] != adj[1]).all())
    key_t = (adj[1] * vs.shape[0] + adj[0]).sort().values
    assert torch.equal(key_t, key)
    # every face references vertices of its own mesh; triangles have side lengths (1, 1, sqrt2)
    counts = torch.tensor(fi, device="cuda")
    vcount = torch.tensor(vi, device="cuda").repeat_interleave(counts)
    assert bool((f >= 0).all()) and bool((f < vcount[:, None]).all())
    voff = (torch.tensor(vi, device="cuda").cumsum(0) - torch.tensor(vi, device="cuda")).repeat_interleave(counts)
    tri = vs[f + voff[:, None]]
    e = torch.stack([(tri[:, 0] - tri[:, 1]).pow(2).sum(1), (tri[:, 1] - tri[:, 2]).pow(2).sum(1),
                     (tri[:, 0] - tri[:, 2]).pow(2).sum(1)], 1).sort(1).values
    assert torch.equal(e, torch.tensor([1.0, 1.0, 2.0], device="cuda").expand_as(e))
    # oracle on the first two meshes (mesh-local outputs are independent of the rest of the batch)
    o = cubify_np.cubify(t[:2].numpy(), 0.5)
    nv, nf = sum(o[1]), sum(o[3])
    assert vi[:2] == o[1] and fi[:2] == o[3]
    assert np.array_equal(vs[:nv].cpu().numpy(), o[0]) and np.array_equal(f[:nf].cpu().numpy(), o[2])
    ne = o[4].shape[1]
    assert np.array_equal(adj[:, :ne].cpu().numpy(), o[4])


def test_cubify_from_logits_and_fused_voxel_loss(lib):
    """the sigmoid that ends the reference's VoxelBranch (layers.py:505) folded into its two consumers.
    Cubify(th)(logits, from_logits=True) == Cubify(th)(sigmoid(logits)) bit for bit (oracle on the CPU probabilities), and
    voxel_loss_with_logits == BCE(sigmoid(logits)) of torch in fp64 (values rtol 1e-6, gradients rtol 1e-4); voxel_loss on
    probabilities likewise (reference loss_functions.py:10-14)."""
    import torch.nn.functional as TF
    from meshrcnn_b200 import loss_functions as LF
    from meshrcnn_b200.layers import Cubify, VoxelBranch
    g = torch.Generator().manual_seed(5)
    B, V, th = 3, 16, 0.2
    logits = torch.randn(B, V, V, V, generator=g) * 3.0
    probs = torch.sigmoid(logits)
    near = (probs - th).abs() < 1e-5                   # keep clear of the threshold: CPU / GPU expf may differ in the last ulp
    logits[near] += 0.01
    probs = torch.sigmoid(logits)
    want = cubify_np.cubify(probs.numpy(), th)
    got = Cubify(th)(logits.cuda(), from_logits=True)
    assert got[1] == want[1] and got[3] == want[3]
    for i in (0, 2, 4):
        assert np.array_equal(got[i].cpu().numpy(), want[i])
    gt = (torch.rand(B, V, V, V, generator=g) < 0.3).float()
    # loss on logits
    x = logits.cuda().requires_grad_()
    loss, p_out = LF.voxel_loss_with_logits(x, gt.cuda(), return_probs=True)
    loss.backward()
    x64 = logits.double().requires_grad_()
    ref = TF.binary_cross_entropy(torch.sigmoid(x64), gt.double(), reduction="mean")
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-6 * abs(float(ref))
    assert torch.allclose(p_out.cpu(), probs, rtol=1e-6, atol=1e-7)
    assert torch.allclose(x.grad.cpu().double(), x64.grad, rtol=1e-4, atol=1e-9)
    # loss on probabilities (the reference signature), incl. saturated entries (log clamp at -100)
    pr = probs.clone()
    pr.view(-1)[:7] = torch.tensor([0.0, 1.0, 1e-30, 1 - 1e-7, 0.5, 0.0, 1.0])
    gt.view(-1)[:7] = torch.tensor([1.0, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0])
    pc = pr.cuda().requires_grad_()
    l2 = LF.voxel_loss(pc, gt.cuda())
    l2.backward()
    p64 = pr.double().requires_grad_()
    r2 = TF.binary_cross_entropy(p64.float(), gt, reduction="mean")
    assert abs(float(l2) - float(r2)) <= 1e-5 * abs(float(r2))
    r64 = TF.binary_cross_entropy(p64, gt.double(), reduction="mean")
    r64.backward()
    inner = (pr > 1e-6) & (pr < 1 - 1e-6)
    assert torch.allclose(pc.grad.cpu().double()[inner], p64.grad[inner], rtol=1e-4, atol=1e-12)
    # the module mirror keeps the reference's state-dict keys and splits off the final sigmoid
    vb = VoxelBranch(8, V, 16).cuda()
    assert list(vb.state_dict()) == ["0.weight", "0.bias", "1.weight", "1.bias", "2.weight", "2.bias", "3.weight", "3.bias"]
    feat = torch.randn(2, 8, V // 2, V // 2, generator=g).cuda()
    assert torch.allclose(torch.sigmoid(vb.forward_logits(feat)), vb(feat), rtol=1e-6, atol=1e-7)

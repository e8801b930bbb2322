"""VertexAlign / GraphConv / refinement stages: CUDA path vs the reference's outputs (tests/golden, produced by
oracle/make_golden.py from the unmodified reference) and vs the oracle on fresh seeded inputs.
Tolerance (north star): fp32 rtol 1e-4 on values and gradients; VertexAlign is a pure gather -> bit-exact."""
import numpy as np
import pytest
import torch

from oracle import cubify_np, mesh_ops

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def cuda(x, grad=False):
    t = torch.as_tensor(np.asarray(x)).cuda()
    if t.is_floating_point():
        t = t.float()
    return t.requires_grad_() if grad else t


def close(got, want, rtol=RTOL, atol=None, what=""):
    want = torch.as_tensor(np.asarray(want)).double()
    got = got.detach().cpu().double()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if atol is None:
        atol = rtol * float(want.abs().max()) * 1e-1 + 1e-30     # gradients span decades: scale atol to the tensor
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    assert bool((err <= tol).all()), "%s: max err %.3e (tol %.3e), max|want| %.3e" % (
        what, float(err.max()), float(tol.min()), float(want.abs().max()))


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,nmaps", [("pix", 1), ("shp", 4), ("randint", 2), ("border", 1)])
def test_vert_align_matches_reference_bit_exact(lib, golden, name, nmaps):
    from meshrcnn_b200.layers import VertexAlign
    g = golden("vert_align")
    hw = int(g[name + "__hw"])
    fm = [cuda(g["%s__fm%d" % (name, i)], grad=True) for i in range(nmaps)]
    pos = cuda(g[name + "__pos"], grad=True)
    vpm = g[name + "__vpm"].tolist()
    out = VertexAlign().eval()(fm, pos, vpm, [(hw, hw)] * 3, [1, 1, 1])
    want = g[name + "_f32__out"]
    assert out.dtype == torch.float32
    assert np.array_equal(out.detach().cpu().numpy(), want), "VertexAlign not bit-exact"
    if name in ("pix", "shp"):
        assert float((torch.as_tensor(want) != 0).float().mean()) > 0.5     # the gather is really exercised
    (out * cuda(g[name + "_f64__gout"])).sum().backward()
    assert pos.grad is None                                                 # reference: no gradient to positions
    for i in range(nmaps):
        close(fm[i].grad, g["%s_f64__gfm%d" % (name, i)], what="gfm%d" % i)


def test_vert_align_mesh_index_and_errors(lib):
    """Several meshes per image (eval path, layers.py:538-543) against the oracle."""
    from meshrcnn_b200 import synthetic
    from meshrcnn_b200.layers import VertexAlign
    fm = synthetic.feature_maps(2, [(8, 12, 12)], 3)
    pos = synthetic.in_frustum_positions(90, 224, 4)
    vpm, mi, sizes = [20, 30, 40], [2, 1], [(224, 224), (224, 224)]
    want = mesh_ops.vert_align(fm, pos, vpm, sizes, mi)
    got = VertexAlign().eval()([f.cuda() for f in fm], pos.cuda(), vpm, sizes, mi)
    assert torch.equal(got.cpu(), want)
    with pytest.raises(RuntimeError):
        VertexAlign().eval()([f.cuda() for f in fm], pos.cuda(), [20, 30], sizes, mi)


def _shapenet_align_case(B=3, n=700, seed=7, scale=0.25):
    from meshrcnn_b200 import synthetic
    shapes = [(c // 8, h, w) for (c, h, w) in synthetic.SHAPENET_MAPS]      # 32 + 64 + 128 + 256 = 480 channels
    fms = [f * scale for f in synthetic.feature_maps(B, shapes, seed)]
    pos = synthetic.in_frustum_positions(n, 137, seed + 1)
    vpm = [n // B] * (B - 1) + [n - (B - 1) * (n // B)]
    g = torch.Generator().manual_seed(seed + 2)
    w = (torch.rand(128, 480, generator=g) - 0.5) * 0.2
    go = torch.randn(n, 128, generator=g)
    return fms, pos, vpm, [(137, 137)] * B, [1] * B, w, go


def test_vert_align_linear_fused_vs_oracle_and_unfused(lib):
    """linear(VertexAlign(maps)) evaluated as per-texel projections + a row gather (csrc/align_proj.cu) against the fp64
    oracle (VertexAlign then the dense bottleneck, reference layers.py:151-155) and against the literal two-step CUDA path:
    values and the gradients of the weight and of every map at rtol 1e-4."""
    from meshrcnn_b200 import functional as F_
    from meshrcnn_b200.layers import VertexAlign
    fms, pos, vpm, sizes, mi, w, go = _shapenet_align_case()

    def oracle(dt):
        f = [x.to(dt).requires_grad_() for x in fms]
        wd = w.to(dt).requires_grad_()
        out = mesh_ops.vert_align(f, pos.to(dt), vpm, sizes, mi) @ wd.t()
        (out * go.to(dt)).sum().backward()
        return out.detach(), wd.grad, [x.grad for x in f]

    want, want_gw, want_gf = oracle(torch.float64)
    assert float((mesh_ops.vert_align(fms, pos, vpm, sizes, mi) != 0).float().mean()) > 0.5    # the gather is exercised

    f_c = [x.cuda().requires_grad_() for x in fms]
    w_c = w.cuda().requires_grad_()
    out = F_.vert_align_linear(f_c, pos.cuda(), vpm, sizes, mi, w_c)
    (out * go.cuda()).sum().backward()
    close(out, want, what="fused out")
    close(w_c.grad, want_gw, what="fused gW")
    for i in range(4):
        close(f_c[i].grad, want_gf[i], what="fused gmap%d" % i)

    f_u = [x.cuda().requires_grad_() for x in fms]
    w_u = w.cuda().requires_grad_()
    out_u = F_.linear(VertexAlign().eval()(f_u, pos.cuda(), vpm, sizes, mi), w_u)
    (out_u * go.cuda()).sum().backward()
    close(out, out_u.detach().cpu(), what="fused vs unfused out")
    close(w_c.grad, w_u.grad.cpu(), what="fused vs unfused gW")


def test_vert_align_linear_all_masked_and_single_map(lib):
    """Pipeline coordinates (half-integer voxel units) all clamp to the image border (SURVEY finding 8): the fused path
    must return exact zeros and zero gradients there; plus the single-map / tiny-row-count (CUDA-core GEMM) branch."""
    from meshrcnn_b200 import functional as F_, synthetic
    fms, _, vpm, sizes, mi, w, go = _shapenet_align_case()
    g = torch.Generator().manual_seed(11)
    n = sum(vpm)                                                  # |p0/p2|, |p1/p2| >= 5: every projection clamps
    pos = torch.stack([10 + 10 * torch.rand(n, generator=g), 10 + 10 * torch.rand(n, generator=g),
                       -(1 + torch.rand(n, generator=g))], 1)
    assert float(mesh_ops.vert_align(fms, pos, vpm, sizes, mi).abs().max()) == 0.0
    f_c = [x.cuda().requires_grad_() for x in fms]
    w_c = w.cuda().requires_grad_()
    out = F_.vert_align_linear(f_c, pos.cuda(), vpm, sizes, mi, w_c)
    assert float(out.abs().max()) == 0.0
    (out * go.cuda()).sum().backward()
    assert float(w_c.grad.abs().max()) == 0.0 and all(float(f.grad.abs().max()) == 0.0 for f in f_c)

    fm = synthetic.feature_maps(2, [(12, 6, 6)], 3)              # K = 12 < 16: CUDA-core fallback GEMMs
    p = synthetic.in_frustum_positions(50, 224, 4)
    g = torch.Generator().manual_seed(5)
    w1 = torch.randn(8, 12, generator=g)
    want = mesh_ops.vert_align([fm[0].double()], p.double(), [20, 30], [(224, 224)] * 2, [1, 1]) @ w1.double().t()
    got = F_.vert_align_linear([fm[0].cuda()], p.cuda(), [20, 30], [(224, 224)] * 2, [1, 1], w1.cuda())
    close(got, want, what="single map")


def test_vert_align_bf16_feature_maps(lib):
    """bf16 feature-map mode (north star: bf16 features rtol 2e-2): the gather of a bf16 map equals the fp32 gather of
    the same (rounded) values exactly, and stays within rtol 2e-2 of the fp64 oracle on the unrounded maps; the fused
    bottleneck path accepts bf16 maps as well.  Gradients come back in bf16."""
    from meshrcnn_b200 import functional as F_
    from meshrcnn_b200.layers import VertexAlign
    fms, pos, vpm, sizes, mi, w, go = _shapenet_align_case()
    want = mesh_ops.vert_align([f.double() for f in fms], pos.double(), vpm, sizes, mi)
    bf = [f.to(torch.bfloat16).cuda().requires_grad_() for f in fms]
    out = VertexAlign().eval()(bf, pos.cuda(), vpm, sizes, mi)
    assert out.dtype == torch.float32
    same = VertexAlign().eval()([b.detach().float() for b in bf], pos.cuda(), vpm, sizes, mi)
    assert torch.equal(out, same)
    close(out, want, rtol=2e-2, what="bf16 gather")
    gsel = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    (out * gsel.cuda()).sum().backward()
    assert all(b.grad is not None and b.grad.dtype == torch.bfloat16 for b in bf)
    f64 = [f.double().requires_grad_() for f in fms]
    (mesh_ops.vert_align(f64, pos.double(), vpm, sizes, mi) * gsel.double()).sum().backward()
    for i in range(4):
        close(bf[i].grad.float(), f64[i].grad, rtol=2e-2, what="bf16 gmap%d" % i)
    # odd channel count -> lane-per-channel kernel
    odd = fms[0][:, :13].contiguous()
    got = VertexAlign().eval()([odd.to(torch.bfloat16).cuda()], pos.cuda(), vpm, sizes, mi)
    close(got, mesh_ops.vert_align([odd.double()], pos.double(), vpm, sizes, mi), rtol=2e-2, what="bf16 odd C")
    # fused bottleneck with bf16 maps
    fused = F_.vert_align_linear([b.detach() for b in bf], pos.cuda(), vpm, sizes, mi, w.cuda())
    close(fused, want @ w.double().t(), rtol=2e-2, what="bf16 fused")


# ---------------------------------------------------------------------------------------------------------------
def test_aggregate_known_answer(lib):
    """reference tests/test_layers.py:16-26 (non-symmetric, unsorted-by-row-safe COO)."""
    from meshrcnn_b200.utils import aggregate_neighbours
    a = torch.tensor([[1., 2, 3], [4, 5, 6], [7, 8, 9]]).cuda()
    ei = torch.tensor([[0, 0, 1, 2], [1, 2, 1, 0]]).cuda()
    out = aggregate_neighbours(ei, a)
    assert torch.allclose(out.cpu(), torch.tensor([[11., 13, 15], [4, 5, 6], [1, 2, 3]]))


def test_aggregate_unsorted_coo_and_grad(lib):
    from meshrcnn_b200.utils import aggregate_neighbours
    g = torch.Generator().manual_seed(0)
    n, E, D = 50, 400, 7
    ei = torch.randint(0, n, (2, E), generator=g)
    m = torch.randn(n, D, generator=g)
    mc = m.cuda().requires_grad_()
    out = aggregate_neighbours(ei.cuda(), mc)
    m64 = m.double().requires_grad_()
    want = mesh_ops.aggregate_neighbours(ei, m64)
    close(out, want.detach(), what="aggregate")
    go = torch.randn(n, D, generator=g)
    (out * go.cuda()).sum().backward()
    (want * go.double()).sum().backward()
    close(mc.grad, m64.grad, what="aggregate grad")


def test_graphconv_known_answer(lib):
    """reference tests/test_layers.py:58-74: all-ones weights on a path graph -> rows [15, 36, 33]."""
    from meshrcnn_b200.layers import GraphConv
    conv = GraphConv(3, 6).cuda()
    with torch.no_grad():
        conv.w0.fill_(1)
        conv.w1.fill_(1)
    x = torch.arange(9).reshape(3, 3).float().cuda()
    adj = torch.tensor([[0, 1, 0], [1, 0, 1], [0, 1, 0]]).nonzero().t().contiguous().cuda()
    out = conv(x, adj)
    assert torch.allclose(out.cpu(), torch.tensor([15., 36, 33]).view(3, 1).expand(3, 6))


@pytest.mark.parametrize("name", ["gc_19_16", "gc_16_3", "gc_35_24"])
def test_graphconv_matches_reference(lib, golden, name):
    from meshrcnn_b200.layers import GraphConv
    g = golden("graphconv_stages")
    w0 = g[name + "__w0"]
    conv = GraphConv(*w0.shape).cuda()
    conv.load_state_dict({"w0": torch.as_tensor(w0), "w1": torch.as_tensor(g[name + "__w1"])})
    x = cuda(g[name + "__x"], grad=True)
    adj = cuda(g["adj"]).long()
    out = conv(x, adj)
    close(out, g[name + "_f64__out"], what="out")
    (out * cuda(g[name + "_f64__gout"])).sum().backward()
    close(x.grad, g[name + "_f64__gx"], what="gx")
    close(conv.w0.grad, g[name + "_f64__gw0"], what="gw0")
    close(conv.w1.grad, g[name + "_f64__gw1"], what="gw1")


@pytest.mark.parametrize("cls", ["ResVertixRefineShapenet", "VertixRefineShapeNet", "VertixRefinePix3D"])
@pytest.mark.parametrize("use_feat", [0, 1])
def test_stage_matches_reference(lib, golden, cls, use_feat):
    """State-dict drop-in: the reference module's own state_dict is loaded into ours."""
    import meshrcnn_b200.layers as L
    g = golden("graphconv_stages")
    tag = "%s_%d" % (cls, use_feat)
    sd = {k[len(tag) + 6:]: torch.as_tensor(v) for k, v in g.items() if k.startswith(tag + "__sd__")}
    nmaps = 1 if cls == "VertixRefinePix3D" else 4
    fm = [cuda(g["%s__fm%d" % (cls, i)], grad=True) for i in range(nmaps)]
    align_c = sum(f.shape[1] for f in fm)
    mod = getattr(L, cls)(use_input_features=bool(use_feat), num_features=16, alignment_size=align_c).cuda().eval()
    mod.load_state_dict(sd, strict=True)
    hw = int(g[cls + "__hw"])
    pos = cuda(g[cls + "__pos"], grad=True)
    feats = cuda(g[cls + "__feats"], grad=True) if use_feat else None
    adj = cuda(g["adj"]).long()
    v_index = g["v_index"].tolist()
    new_pos, new_feat = mod(v_index, fm[0] if nmaps == 1 else fm, adj, pos, [(hw, hw)] * 2, vertex_features=feats)
    t64 = tag + "_f64"
    close(new_pos, g[t64 + "__new_pos"], what="new_pos")
    close(new_feat, g[t64 + "__new_feat"], what="new_feat")
    ((new_pos * cuda(g[tag + "_f64__gpos_out"])).sum() + (new_feat * cuda(g[tag + "_f64__gfeat_out"])).sum()).backward()
    close(pos.grad, g[t64 + "__grad__pos"], what="grad pos")
    for i in range(nmaps):
        close(fm[i].grad, g[t64 + "__grad__fm%d" % i], what="grad fm%d" % i)
    if use_feat:
        close(feats.grad, g[t64 + "__grad__feats"], what="grad feats")
    for k, p in mod.named_parameters():
        close(p.grad, g[t64 + "__grad__param__" + k], what="grad " + k)


def test_stage_chain_on_cubified_batch_vs_oracle(lib):
    """Cubify -> 3 Pix3D stages at production width (128 features, 256-channel 12x12 map), fwd + bwd vs the oracle in
    fp64 (positions are in-frustum so the gather is exercised; the graph is the cubified one)."""
    import meshrcnn_b200.layers as L
    from meshrcnn_b200 import synthetic
    B, V = 3, 12
    vox = synthetic.blob_voxels(B, V, 0)
    verts, v_index, faces, f_index, adj = L.Cubify(0.2)(vox.cuda())
    o = cubify_np.cubify(vox.numpy(), 0.2)
    assert np.array_equal(adj.cpu().numpy(), o[4])
    SV = verts.shape[0]
    torch.manual_seed(1)
    stages = [L.VertixRefinePix3D(use_input_features=bool(i)).cuda() for i in range(3)]
    with torch.no_grad():                      # default init + neighbour sums grow ~20x per layer and saturate the tanh
        for st in stages:
            for prm in st.parameters():
                prm.mul_(0.2)
    fmap = synthetic.feature_maps(B, [synthetic.PIX3D_MAP], 0)[0] * 0.02     # keeps the tanh heads unsaturated
    pos0 = synthetic.in_frustum_positions(SV, 224, 5)
    sizes = [(224, 224)] * B

    fm_c = fmap.cuda().requires_grad_()
    p = pos0.cuda().requires_grad_()
    feats = None
    cur = p
    for st in stages:
        cur, feats = st(v_index, fm_c, adj, cur, sizes, vertex_features=feats)
    go_p = torch.randn(SV, 3, generator=torch.Generator().manual_seed(2))
    (cur * go_p.cuda()).sum().backward()

    def oracle(dt):
        fm = fmap.to(dt).requires_grad_()
        p_ = pos0.to(dt).requires_grad_()
        params = [{k: v.detach().cpu().to(dt).requires_grad_() for k, v in st.named_parameters()} for st in stages]
        c, f = p_, None
        for sd in params:
            c, f = mesh_ops.stage_pix3d(sd, v_index, fm, adj.cpu(), c, sizes, feats=f)
        (c * go_p.to(dt)).sum().backward()
        out = {"pos3": c.detach(), "feat3": f.detach(), "grad pos0": p_.grad, "grad fmap": fm.grad}
        for i, sd in enumerate(params):
            for k, v in sd.items():
                out["grad %d.%s" % (i, k)] = v.grad
        return out

    o64, o32 = oracle(torch.float64), oracle(torch.float32)
    got = {"pos3": cur, "feat3": feats, "grad pos0": p.grad, "grad fmap": fm_c.grad}
    for i, st in enumerate(stages):
        for k, prm in st.named_parameters():
            got["grad %d.%s" % (i, k)] = prm.grad
    # Deep fp32 chain: ReLU masks of pre-activations within one ulp of zero differ between any two fp32 evaluations, so
    # the bound is rtol 1e-4 in relative L2 norm OR 3x the reference's own |fp32 - fp64| deviation, whichever is larger.
    for name, want in o64.items():
        w = want.double()
        assert float(w.norm()) > 0, name
        err = float((got[name].detach().cpu().double() - w).norm() / w.norm())
        ref = float((o32[name].double() - w).norm() / w.norm())
        # Values: rtol 1e-4 (relative L2).  Gradients of a deep chain: any two fp32 evaluations differ in the ReLU mask of
        # the few pre-activations within ~1e-6 of zero (3xTF32 products are accurate to 6e-7, IEEE fp32 to 3e-7); one
        # flipped mask moves a weight gradient summed over ~500 vertices by ~1e-4.  Per-layer gradient parity at rtol 1e-4
        # is tested on identical inputs above; the chain is bounded by max(1e-4, 3x the reference's own |fp32 - fp64|),
        # and 1e-3 for gradients.
        bound = max(1e-4, 3 * ref)
        if name.startswith("grad"):
            bound = max(bound, 1e-3)
        assert err <= bound, (name, err, ref)


@pytest.mark.parametrize("model", ["shapenet_residual", "shapenet"])
def test_shapenet_heads_production_widths_vs_oracle(lib, model):
    """BASELINE config-1 style ShapeNet heads at production widths (4 maps = 3840 channels for 137-px images, 128
    features, the 3840 -> 128 bottleneck on the multi-accumulator tensor-core path): stage chain fwd + bwd vs the oracle."""
    from meshrcnn_b200 import synthetic
    from meshrcnn_b200.pipeline import RefinementHead
    B, V = 2, 10
    vox = synthetic.blob_voxels(B, V, 3)
    torch.manual_seed(2)
    head = RefinementHead(model, cubify_threshold=0.2).cuda().eval()
    with torch.no_grad():
        for prm in head.parameters():
            prm.mul_(0.2)
    fmaps = [m * 0.05 for m in synthetic.feature_maps(B, synthetic.SHAPENET_MAPS, 1)]
    verts, v_index, faces, f_index, adj = head.cubify(vox.cuda())
    SV = verts.shape[0]
    pos0 = synthetic.in_frustum_positions(SV, 137, 9)
    sizes = [(137, 137)] * B
    fm_c = [m.cuda().requires_grad_() for m in fmaps]
    p = pos0.cuda().requires_grad_()
    cur, feats = p, None
    for st in head.refineStages:
        cur, feats = st(v_index, fm_c, adj, cur, sizes, vertex_features=feats)
    go = torch.randn(SV, 3, generator=torch.Generator().manual_seed(4))
    (cur * go.cuda()).sum().backward()

    stage_fn = mesh_ops.STAGES[type(head.refineStages[0]).__name__]

    def oracle(dt):
        fm = [m.to(dt).requires_grad_() for m in fmaps]
        p_ = pos0.to(dt).requires_grad_()
        params = [{k: v.detach().cpu().to(dt).requires_grad_() for k, v in st.named_parameters()} for st in head.refineStages]
        c, f = p_, None
        for sd in params:
            c, f = stage_fn(sd, v_index, fm, adj.cpu(), c, sizes, feats=f)
        (c * go.to(dt)).sum().backward()
        out = {"pos3": c.detach(), "feat3": f.detach(), "grad pos0": p_.grad}
        for i, m in enumerate(fm):
            out["grad fmap%d" % i] = m.grad
        for i, sd in enumerate(params):
            for k, v in sd.items():
                out["grad %d.%s" % (i, k)] = v.grad
        return out

    o64, o32 = oracle(torch.float64), oracle(torch.float32)
    got = {"pos3": cur, "feat3": feats, "grad pos0": p.grad}
    for i, m in enumerate(fm_c):
        got["grad fmap%d" % i] = m.grad
    for i, st in enumerate(head.refineStages):
        for k, prm in st.named_parameters():
            got["grad %d.%s" % (i, k)] = prm.grad
    for name, want in o64.items():
        w = want.double()
        assert float(w.norm()) > 0, name
        err = float((got[name].detach().cpu().double() - w).norm() / w.norm())
        ref = float((o32[name].double() - w).norm() / w.norm())
        # Values: rtol 1e-4 (relative L2).  Gradients of a deep chain: any two fp32 evaluations differ in the ReLU mask of
        # the few pre-activations within ~1e-6 of zero (3xTF32 products are accurate to 6e-7, IEEE fp32 to 3e-7); one
        # flipped mask moves a weight gradient summed over ~500 vertices by ~1e-4.  Per-layer gradient parity at rtol 1e-4
        # is tested on identical inputs above; the chain is bounded by max(1e-4, 3x the reference's own |fp32 - fp64|),
        # and 1e-3 for gradients.
        bound = max(1e-4, 3 * ref)
        if name.startswith("grad"):
            bound = max(bound, 1e-3)
        assert err <= bound, (name, err, ref)


def test_concat_cols_matches_torch_cat_with_strided_parts(lib):
    """``concat_cols`` (16-byte aligned rows, gradient slices as views) vs ``torch.cat`` incl. strided inputs and the
    GraphConv that consumes the padded view (reference layers.py:160-165,241-252,321-334)."""
    from meshrcnn_b200 import functional as F_
    g = torch.Generator().manual_seed(5)
    n = 777
    pos = torch.randn(n, 3, generator=g).cuda().requires_grad_()
    wide = torch.randn(n, 140, generator=g).cuda().requires_grad_()
    x = wide[:, 5:133]                                   # 128-wide strided view, not 16-byte aligned
    feat = torch.randn(n, 256, generator=g).cuda().requires_grad_()
    got = F_.concat_cols([x, pos, feat])
    want = torch.cat([x, pos, feat], 1)
    assert got.shape == want.shape and got.stride(0) % 4 == 0 and torch.equal(got, want)
    w = torch.randn(n, 387, generator=g).cuda()
    (got * w).sum().backward()
    gpos, gwide, gfeat = pos.grad.clone(), wide.grad.clone(), feat.grad.clone()
    pos.grad = wide.grad = feat.grad = None
    (want * w).sum().backward()
    assert torch.equal(gpos, pos.grad) and torch.equal(gwide, wide.grad) and torch.equal(gfeat, feat.grad)
    # GraphConv on the padded view == GraphConv on a contiguous copy
    adj = torch.stack([torch.arange(n), (torch.arange(n) + 1) % n]).cuda()
    adj = torch.cat([adj, adj.flip(0)], 1)
    adj = adj[:, torch.argsort(adj[0] * n + adj[1])]
    w0, w1 = (torch.randn(387, 128, generator=g) * 0.05).cuda().requires_grad_(), (torch.randn(387, 128, generator=g) * 0.05).cuda().requires_grad_()
    outs = []
    for inp in (got.detach().requires_grad_(), want.detach().contiguous().requires_grad_()):
        w0.grad = w1.grad = None
        y = F_.graph_conv(inp, adj, w0, w1)
        y.square().sum().backward()
        outs.append((y.detach(), inp.grad.clone(), w0.grad.clone(), w1.grad.clone()))
    for a, b, what in zip(outs[0], outs[1], ("out", "gx", "gw0", "gw1")):
        close(a, b.cpu(), what=what)


def test_overlapped_losses_equal_single_stream(lib):
    """The per-stage losses issued on the second CUDA stream (pipeline.RefinementHead.overlap_losses) give the same losses
    and gradients as the single-stream order of the reference (shapenet_model.py:92-95)."""
    from meshrcnn_b200 import synthetic
    from meshrcnn_b200.layers import Cubify
    from meshrcnn_b200.mesh_sampling import normalize_mesh
    from meshrcnn_b200.pipeline import MeshTargets, RefinementHead, weighted_loss
    B, n = 3, 2000
    vox = synthetic.blob_voxels(B, 16, 3).cuda()
    fmap = synthetic.feature_maps(B, [synthetic.PIX3D_MAP], 3)[0].cuda().requires_grad_()
    gv, gvi, gf, gfi, _ = Cubify(0.5)(synthetic.blob_voxels(B, 16, 1003).cuda())
    gt = MeshTargets(torch.cat([normalize_mesh(v) for v in gv.split(gvi)]), gf, gvi, gfi)
    torch.manual_seed(2)
    head = RefinementHead("pix3d", cubify_threshold=0.2).cuda().train()
    rnd = []
    for s in range(3):
        u, x2, x1 = synthetic.sampling_randomness(B, 10000, 50 + s)
        ug, x2g, x1g = synthetic.sampling_randomness(B, 10000, 60 + s)
        rnd.append(({"u": u.cuda(), "xi2": x2.cuda(), "xi1": x1.cuda()}, {"u": ug.cuda(), "xi2": x2g.cuda(), "xi1": x1g.cuda()}))
    res = []
    for overlap in (True, False):
        head.overlap_losses = overlap
        head.zero_grad(set_to_none=True)
        fmap.grad = None
        losses = head(vox, fmap, [(224, 224)] * B, gt, loss_randomness=rnd)
        weighted_loss(losses).backward()
        torch.cuda.synchronize()
        res.append(({k: float(v) for k, v in losses.items()}, fmap.grad.clone(), [p.grad.clone() for p in head.parameters()]))
    assert res[0][0].keys() == res[1][0].keys()
    for k in res[0][0]:
        assert abs(res[0][0][k] - res[1][0][k]) <= 1e-5 * abs(res[1][0][k]) + 1e-7, (k, res[0][0][k], res[1][0][k])
    close(res[0][1], res[1][1].cpu(), what="fmap grad")
    for a, b in zip(res[0][2], res[1][2]):
        close(a, b.cpu(), what="param grad")


def test_eval_mode_head_output_dict_vs_oracle(lib):
    """The eval-mode output dict of the refinement-stage loop (reference shapenet_model.py:96-99 / pix3d_model.py:112-115):
    keys, list of 4 position sets, Cubify topology -- final positions against the fp64 oracle chain; and
    ``sharding.gather_eval_outputs`` on the real output (single process: returned unchanged)."""
    from meshrcnn_b200 import synthetic
    from meshrcnn_b200.pipeline import RefinementHead
    from meshrcnn_b200.sharding import gather_eval_outputs
    B, V = 3, 12
    vox = synthetic.blob_voxels(B, V, 5)
    fmap = synthetic.feature_maps(B, [synthetic.PIX3D_MAP], 5)[0] * 0.02
    sizes = [(224, 224)] * B
    torch.manual_seed(4)
    head = RefinementHead("pix3d", cubify_threshold=0.2).cuda().eval()
    with torch.no_grad():
        for prm in head.parameters():
            prm.mul_(0.2)
        out = head(vox.cuda(), fmap.cuda(), sizes)
    assert set(out) == {"vertex_positions", "edge_index", "face_index", "vertice_index", "faces", "mesh_index"}
    o = cubify_np.cubify(vox.numpy(), 0.2)
    assert out["vertice_index"] == o[1] and out["face_index"] == o[3] and out["mesh_index"] == [1] * B
    assert np.array_equal(out["faces"].cpu().numpy(), o[2]) and np.array_equal(out["edge_index"].cpu().numpy(), o[4])
    assert len(out["vertex_positions"]) == 4 and np.array_equal(out["vertex_positions"][0].cpu().numpy(), o[0])
    cur, feats = torch.from_numpy(o[0]).double(), None
    for s, st in enumerate(head.refineStages):
        sd = {k: v.detach().cpu().double() for k, v in st.named_parameters()}
        cur, feats = mesh_ops.stage_pix3d(sd, o[1], fmap.double(), torch.from_numpy(o[4]), cur, sizes, feats=feats)
        close(out["vertex_positions"][s + 1], cur, what="eval positions of stage %d" % s)
    assert gather_eval_outputs(out) is out
    with pytest.raises(ValueError):
        head.train()(vox.cuda(), fmap.cuda(), sizes)                 # training mode needs targets (shapenet_model.py:58-59)


def test_two_devices_two_threads_like_reference_dp(lib):
    """The reference's CustomDP drives one replica per GPU from Python threads of ONE process (dataParallel.py:33).  The
    > 48 KB shared-memory opt-in is per device: both devices must be able to run the tcgen05 GEMM and the cell-grid k-NN
    concurrently, and a tensor on the wrong device must be rejected."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import threading
    from meshrcnn_b200 import functional as F_
    res, err = {}, []

    def work(i):
        try:
            with torch.cuda.device(i):
                g = torch.Generator().manual_seed(i)
                x = torch.randn(3000, 131, generator=g)
                w = torch.randn(131, 256, generator=g)
                y = F_.matmul(x.cuda(i), w.cuda(i))
                p, q = torch.rand(2, 3000, 3, generator=g), torch.rand(2, 3000, 3, generator=g)
                d = F_.knn_search(p.cuda(i), q.cuda(i), 10)
                torch.cuda.synchronize(i)
                res[i] = (float((y.cpu().double() - x.double() @ w.double()).abs().max()),
                          bool(torch.equal(d[1].cpu().long(), mesh_ops.p2p_distance(p.double(), q.double()).argmin(2))))
        except Exception as e:      # noqa: BLE001
            err.append((i, repr(e)))

    for rep in range(2):
        ts = [threading.Thread(target=work, args=(i,)) for i in (1, 0)]        # device 1 first: the opt-in is not "done" by dev 0
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not err, err
        assert all(res[i][0] < 1e-3 and res[i][1] for i in (0, 1)), res
    with torch.cuda.device(0), pytest.raises(RuntimeError):
        F_.matmul(torch.randn(64, 32).cuda(1), torch.randn(32, 32).cuda(1))


# ---------------------------------------------------------------------------------------------------------------
# split-input GraphConv (csrc/graphconv2.cu): parts vs the literal concatenation vs the fp64 oracle
# ---------------------------------------------------------------------------------------------------------------
def _parts_case(with_feat, with_tex, D=32, seed=3):
    from meshrcnn_b200 import synthetic
    import meshrcnn_b200.layers as L
    B, V = 3, 10
    vox = synthetic.blob_voxels(B, V, seed)
    verts, v_index, faces, f_index, adj = L.Cubify(0.2)(vox.cuda())
    SV = verts.shape[0]
    g = torch.Generator().manual_seed(seed)
    pos = synthetic.in_frustum_positions(SV, 224, seed)
    feat = torch.randn(SV, 48, generator=g) if with_feat else None
    fmap = synthetic.feature_maps(B, [(20, 12, 12)], seed)[0] if with_tex else None
    K = (48 if with_feat else 0) + 3 + (20 if with_tex else 0)
    w0, w1 = torch.randn(K, D, generator=g) * 0.3, torch.randn(K, D, generator=g) * 0.3
    res = torch.randn(SV, D, generator=g)
    go = torch.randn(SV, D, generator=g)
    return dict(v_index=v_index, adj=adj, SV=SV, pos=pos, feat=feat, fmap=fmap, w0=w0, w1=w1, res=res, go=go, B=B)


@pytest.mark.parametrize("with_feat,with_tex,with_res", [(True, True, False), (False, True, True), (True, False, True),
                                                         (False, False, False)])
def test_graph_conv_parts_vs_oracle(lib, with_feat, with_tex, with_res):
    """relu([feat | pos | aligned] W0 + A([..] W1)) (+ residual) evaluated from its parts: values and every gradient
    (parts, feature map, weights, residual) against the fp64 oracle on the materialised concatenation."""
    from meshrcnn_b200 import functional as F_
    c = _parts_case(with_feat, with_tex)
    sizes, mi = [(224, 224)] * c["B"], [1] * c["B"]

    def oracle(dt):
        pos = c["pos"].to(dt).requires_grad_()
        feat = c["feat"].to(dt).requires_grad_() if with_feat else None
        fmap = c["fmap"].to(dt).requires_grad_() if with_tex else None
        w0, w1 = c["w0"].to(dt).requires_grad_(), c["w1"].to(dt).requires_grad_()
        res = c["res"].to(dt).requires_grad_() if with_res else None
        cols = ([feat] if with_feat else []) + [pos] + ([mesh_ops.vert_align([fmap], pos.detach(), c["v_index"], sizes, mi)] if with_tex else [])
        out = mesh_ops.graph_conv(torch.cat(cols, 1), c["adj"].cpu(), w0, w1)
        if with_res:
            out = out + res
        (out * c["go"].to(dt)).sum().backward()
        return dict(out=out.detach(), pos=pos.grad, feat=None if feat is None else feat.grad, fmap=None if fmap is None else fmap.grad,
                    w0=w0.grad, w1=w1.grad, res=None if res is None else res.grad)

    want = oracle(torch.float64)
    pos = c["pos"].cuda().requires_grad_()
    feat = c["feat"].cuda().requires_grad_() if with_feat else None
    fmap = c["fmap"].cuda().requires_grad_() if with_tex else None
    w0, w1 = c["w0"].cuda().requires_grad_(), c["w1"].cuda().requires_grad_()
    res = c["res"].cuda().requires_grad_() if with_res else None
    parts = ([("x", feat)] if with_feat else []) + [("pos", pos)]
    if with_tex:
        tex = F_.TexelTerm(fmap, pos, c["v_index"], sizes, mi)
        assert float((tex.texrow >= 0).float().mean()) > 0.5            # the gather is exercised
        parts.append(("tex", tex))
    out = F_.graph_conv_parts(parts, c["adj"], w0, w1, res)
    (out * c["go"].cuda()).sum().backward()
    got = dict(out=out, pos=pos.grad, feat=None if feat is None else feat.grad, fmap=None if fmap is None else fmap.grad,
               w0=w0.grad, w1=w1.grad, res=None if res is None else res.grad)
    for k, wv in want.items():
        if wv is None:
            continue
        close(got[k], wv, what="parts %s" % k)


def test_pack_plan_batches_weight_images_without_changing_results(lib):
    """functional.PackPlan: the second pass through the same blocks packs all weight images with ONE launch, re-packs after
    a weight update, and values / gradients are bit-identical to the per-block path."""
    from meshrcnn_b200 import functional as F_, _lib
    c = _parts_case(True, True)
    sizes, mi = [(224, 224)] * c["B"], [1] * c["B"]
    w0, w1 = c["w0"].cuda().requires_grad_(), c["w1"].cuda().requires_grad_()

    def run(plan):
        pos, feat, fmap = c["pos"].cuda().requires_grad_(), c["feat"].cuda().requires_grad_(), c["fmap"].cuda().requires_grad_()
        w0.grad = w1.grad = None
        calls = []
        orig = _lib.call
        _lib.call = lambda name, *a: (calls.append(name), orig(name, *a))[1]
        try:
            with F_.pack_plan(plan, fmap) as active:    # with the map: its texel rows can be projected ahead of the mesh
                if active is not None:
                    active.launch_early()               # (Cubify calls this between its count kernels and their read-back)
                tex = F_.TexelTerm(fmap, pos, c["v_index"], sizes, mi)
                out = F_.graph_conv_parts([("x", feat), ("pos", pos), ("tex", tex)], c["adj"], w0, w1, None)
                (out * c["go"].cuda()).sum().backward()
        finally:
            _lib.call = orig
        return [out.detach(), pos.grad, feat.grad, fmap.grad, w0.grad.clone(), w1.grad.clone()], calls

    ref, calls0 = run(None)
    plan = F_.PackPlan()
    first, calls1 = run(plan)                      # nothing on the list yet: per-block packing, the blocks get recorded
    second, calls2 = run(plan)                     # one batched launch
    packs = lambda cs: [n for n in cs if "pack" in n]
    assert packs(calls1) == packs(calls0) and len(packs(calls0)) == 2
    assert packs(calls2) == ["mrb_gemm_tc_pack_graphconv_batch"]
    first_gemm = lambda cs: [n for n in cs if n.startswith("mrb_gemm_tc_acc") or n.startswith("mrb_gc_") or "texrows" in n][0]
    assert first_gemm(calls1) == "mrb_vert_align_texrows"
    if F_.EARLY_TEXEL_PROJECTION:                  # (MRB_EARLY_TEXEL=0 switches the hoisting off)
        assert first_gemm(calls2) == "mrb_gemm_tc_acc"                                                  # texel projection hoisted
    assert calls2.count("mrb_gemm_tc_acc") == calls1.count("mrb_gemm_tc_acc")                           # ... not duplicated
    assert torch.equal(ref[0], first[0]) and torch.equal(ref[0], second[0])          # forward values: bit-identical
    for a, b, c_ in zip(ref[1:], first[1:], second[1:]):                             # gradients: atomic summation order only
        scale = float(a.abs().max())
        assert float((a - b).abs().max()) <= 2e-5 * scale and float((a - c_).abs().max()) <= 2e-5 * scale
    with torch.no_grad():                          # an optimiser step: the next pass must pack the NEW weights
        w0.mul_(0.5)
    third, calls3 = run(plan)
    fresh, _ = run(None)
    assert packs(calls3) == ["mrb_gemm_tc_pack_graphconv_batch"]
    assert torch.equal(third[0], fresh[0]) and not torch.equal(third[0], ref[0])


def test_position_head_vs_oracle(lib):
    """pos + tanh([pos | x] W^T) (Pix3D), pos + tanh(x W^T) (ShapeNet): one kernel forward, fused backward."""
    from meshrcnn_b200 import functional as F_
    g = torch.Generator().manual_seed(9)
    n, Kx = 1001, 128
    x, pos = torch.randn(n, Kx, generator=g) * 0.3, torch.randn(n, 3, generator=g)
    go = torch.randn(n, 3, generator=g)
    for pos_first in (True, None):
        w = torch.randn(3, Kx + (3 if pos_first else 0), generator=g) * 0.1
        xd, pd, wd = (t.double().requires_grad_() for t in (x, pos, w))
        inp = torch.cat([pd, xd], 1) if pos_first else xd
        want = pd + torch.tanh(inp @ wd.t())
        (want * go.double()).sum().backward()
        xc, pc, wc = (t.cuda().requires_grad_() for t in (x, pos, w))
        got = F_.position_head(xc, pc, wc, pos_first)
        (got * go.cuda()).sum().backward()
        close(got, want.detach(), what="head out")
        close(xc.grad, xd.grad, what="head gx")
        close(pc.grad, pd.grad, what="head gpos")
        close(wc.grad, wd.grad, what="head gW")


@pytest.mark.parametrize("model", ["pix3d", "shapenet", "shapenet_residual"])
def test_fused_stages_equal_literal_stages(lib, model):
    """FUSE_STAGE_INPUTS (parts, texel gather, fused head, residual epilogue) against the literal evaluation (concatenate,
    generic kernels) of the same stage chain: outputs and all parameter / map gradients at rtol 1e-4 (relative L2)."""
    import meshrcnn_b200.layers as L
    from meshrcnn_b200 import synthetic
    from meshrcnn_b200.pipeline import RefinementHead
    B, V = 2, 10
    vox = synthetic.blob_voxels(B, V, 3)
    torch.manual_seed(2)
    head = RefinementHead(model, cubify_threshold=0.2).cuda().eval()
    with torch.no_grad():
        for prm in head.parameters():
            prm.mul_(0.2)
    img = 224 if model == "pix3d" else 137
    maps = [synthetic.PIX3D_MAP] if model == "pix3d" else synthetic.SHAPENET_MAPS
    fmaps = [m * 0.05 for m in synthetic.feature_maps(B, maps, 1)]
    verts, v_index, faces, f_index, adj = head.cubify(vox.cuda())
    SV = verts.shape[0]
    pos0 = synthetic.in_frustum_positions(SV, img, 9)
    sizes = [(img, img)] * B
    go = torch.randn(SV, 3, generator=torch.Generator().manual_seed(4)).cuda()
    gf = torch.randn(SV, 128, generator=torch.Generator().manual_seed(5)).cuda()

    def run(fused):
        L.FUSE_STAGE_INPUTS = fused
        try:
            head.zero_grad(set_to_none=True)
            fm = [m.cuda().requires_grad_() for m in fmaps]
            p = pos0.cuda().requires_grad_()
            cur, feats = p, None
            for st in head.refineStages:
                cur, feats = st(v_index, fm[0] if model == "pix3d" else fm, adj, cur, sizes, vertex_features=feats)
            ((cur * go).sum() + (feats * gf).sum()).backward()
            out = {"pos": cur.detach(), "feat": feats.detach(), "g pos0": p.grad}
            for i, m in enumerate(fm):
                out["g fmap%d" % i] = m.grad
            for k, prm in head.named_parameters():
                out["g " + k] = prm.grad.clone()
            return out
        finally:
            L.FUSE_STAGE_INPUTS = True

    a, b = run(True), run(False)
    assert a.keys() == b.keys()
    for k in a:
        w = b[k].double()
        err = float((a[k].double() - w).norm() / max(float(w.norm()), 1e-30))
        assert err <= 1e-4, (k, err)


def test_foreign_coo_validation_and_topology_cache(lib):
    """out-of-range vertex ids in a foreign edge list raise IndexError (like the reference's indexing) instead of
    being silently rewritten, and an in-place edit of a cached edge list invalidates its CSR."""
    from meshrcnn_b200 import functional as F_, topology
    x = torch.randn(6, 8).cuda()
    good = torch.tensor([[0, 1, 2, 3, 4, 5], [1, 0, 3, 2, 5, 4]]).cuda()
    out = F_.aggregate_neighbours(good, x)
    assert torch.equal(out.cpu(), x.cpu()[[1, 0, 3, 2, 5, 4]])
    for bad in (torch.tensor([[0, 1, 7], [1, 0, 2]]), torch.tensor([[0, 1, 2], [1, 0, 6]]), torch.tensor([[0, -1], [1, 0]])):
        with pytest.raises(IndexError):
            F_.aggregate_neighbours(bad.cuda(), x)
    t0 = topology.from_coo(good, 6)
    assert topology.from_coo(good, 6) is t0                     # cached
    good[1, 0] = 2                                              # in-place edit: edge (0,1) -> (0,2)
    t1 = topology.from_coo(good, 6)
    assert t1 is not t0
    want = x.cpu()[[2, 0, 3, 2, 5, 4]]
    assert torch.equal(F_.aggregate_neighbours(good, x).cpu(), want)
    empty = torch.zeros(2, 0, dtype=torch.int64).cuda()
    assert float(F_.aggregate_neighbours(empty, x).abs().sum()) == 0.0

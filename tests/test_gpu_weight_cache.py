"""The prepared-weight cache (sparseconvnet/ops.py _WeightCache, b200scn_prep_weight_tf32_batch): operands are refreshed once
per optimiser step for all layers in one launch and must always reflect the current parameters."""
import pytest
import torch

from _util import random_cloud, rel_err

pytestmark = pytest.mark.gpu


def _net(scn):
    return scn.Sequential(scn.InputLayer(3, 4096, mode=4), scn.SubmanifoldConvolution(3, 3, 16, 3, False),
                          scn.BatchNormReLU(16), scn.SubmanifoldConvolution(3, 16, 32, 3, False),
                          scn.Convolution(3, 32, 32, 2, 2, False), scn.NetworkInNetwork(32, 16, False),
                          scn.SubmanifoldConvolution(3, 16, 16, 3, False), scn.UnPooling(3, 2, 2), scn.OutputLayer(3))


def _same(a, b):
    """Equal up to fp32 summation order: small levels split their offsets over CTAs that add with fp32 atomics and the
    BatchNorm statistics are summed with atomics too -- measured run-to-run spread of this net, cache on or off: 1-2e-5."""
    return rel_err(a, b) < 1e-4


def _run(net, coords, feats):
    out = net([coords, feats])
    out.pow(2).mean().backward()
    return out.detach().clone(), [p.grad.detach().clone() for p in net.parameters()]


def test_cached_operands_follow_the_parameters():
    import sparseconvnet as scn
    from sparseconvnet import ops
    scn.set_precision("tf32")
    try:
        coords, feats = random_cloud(11, 3000, 20, 2, dup_frac=0.3)
        feats = feats.cuda()
        torch.manual_seed(0)
        net = _net(scn).cuda()
        opt = torch.optim.SGD(net.parameters(), lr=0.05)
        scn.set_weight_cache(True)
        for step in range(3):                       # forward/backward, optimiser step, repeat: weights change every step
            opt.zero_grad()
            before = scn.launch_count()
            out_on, _ = _run(net, coords, feats)
            launches = scn.launch_count() - before
            if step == 0:
                first_launches = launches
            # the same weights through one preparation launch per layer and call (cache bypassed, its state untouched)
            ops._weight_cache_on[0] = False
            with torch.no_grad():
                out_off = net([coords, feats])
            ops._weight_cache_on[0] = True
            assert _same(out_on, out_off), step
            opt.step()
        # steps after the first prepare all layers with ONE launch instead of one per layer (4 cached layers here)
        assert launches == first_launches - 4 + 1, (first_launches, launches)
        # parameters replaced wholesale (checkpoint round trip through the CPU, load_state_dict): the cache follows
        scn.set_weight_cache(True)
        out_a, _ = _run(net, coords, feats)
        net.cpu(); net.cuda()
        out_b, _ = _run(net, coords, feats)
        assert _same(out_a, out_b)
        sd = {k: v * 0.5 if v.dtype == torch.float32 and "running" not in k else v for k, v in net.state_dict().items()}
        net.load_state_dict(sd)
        out_c, _ = _run(net, coords, feats)
        scn.set_weight_cache(False)
        out_d, _ = _run(net, coords, feats)
        assert _same(out_c, out_d) and rel_err(out_c, out_b) > 1e-3
        # an edit through .data is invisible to the version counter: invalidate_weight_cache() is the documented remedy
        scn.set_weight_cache(True)
        _run(net, coords, feats)
        for p in net.parameters():
            p.data.mul_(2.0)
        scn.invalidate_weight_cache()
        out_e, _ = _run(net, coords, feats)
        scn.set_weight_cache(False)
        out_f, _ = _run(net, coords, feats)
        assert _same(out_e, out_f)
    finally:
        scn.set_weight_cache(True)
        scn.set_precision("fp32")

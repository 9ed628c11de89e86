"""Host-side logic of the scn surface that needs no GPU: builders, parameter shapes, checkpoints, counters, weights."""
import os

import numpy as np
import pytest
import torch

import sparseconvnet as scn
from b200scn_synth import build_encoder, make_batch, make_scene
from sparseconvnet.ops import GemmWeight


def _nparams(net):
    return sum(p.numel() for p in net.parameters())


def test_encoder_parameter_counts_match_survey_appendix_c():
    assert _nparams(build_encoder(scn, "SparseConvUNet", 16, 1, False)) == 2689520
    assert _nparams(build_encoder(scn, "SparseConvUNet", 32, 2, True)) == 30103712
    assert _nparams(build_encoder(scn, "SparseConvFCNet", 16, 1, False)) == 1200816
    assert _nparams(build_encoder(scn, "SparseConvFCNet", 16, 2, True)) == 4106544


def test_state_dict_layout_is_scn_compatible():
    from oracle import scn_oracle as ref
    a = build_encoder(scn, "SparseConvUNet", 16, 1, False)
    b = build_encoder(ref, "SparseConvUNet", 16, 1, False)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    assert all(sa[k].shape == sb[k].shape for k in sa)
    conv = scn.SubmanifoldConvolution(3, 16, 32, 3, False)
    assert conv.weight.shape == (27, 1, 16, 32)                # (filter_volume, groups, nIn, nOut)
    assert scn.Convolution(3, 16, 32, 2, 2, False).weight.shape == (8, 1, 16, 32)
    assert scn.Convolution(3, 16, 32, 4, 4, False).weight.shape == (64, 1, 16, 32)
    assert scn.NetworkInNetwork(16, 32, False).weight.shape == (16, 32)
    bn = scn.BatchNormReLU(8)
    assert set(bn.state_dict()) == {"weight", "bias", "running_mean", "running_var"}
    assert bn.eps == 1e-4 and bn.momentum == 0.9 and bn.leakiness == 0
    # Lua-era camelCase buffer names load too
    sd = {"weight": torch.ones(8), "bias": torch.zeros(8), "runningMean": torch.full((8,), 2.0), "runningVar": torch.ones(8)}
    bn.load_state_dict(sd)
    assert float(bn.running_mean[0]) == 2.0


def test_constructor_errors_mirror_scn_limits():
    with pytest.raises(NotImplementedError):
        scn.SubmanifoldConvolution(3, 4, 4, 5, False)
    with pytest.raises(NotImplementedError):
        scn.Convolution(3, 4, 4, 3, 2, False)
    with pytest.raises(NotImplementedError):
        scn.InputLayer(2, 64, mode=4)
    m = scn.Sequential().add(scn.Identity()).add(scn.Convolution(3, 4, 8, 2, 2, False))
    assert len(m) == 2 and m.input_spatial_size(torch.LongTensor([8])).item() == 16


def test_checkpoint_roundtrip(tmp_path):
    net = build_encoder(scn, "SparseConvUNet", 16, 1, False)
    exp = os.path.join(tmp_path, "exp")
    assert scn.checkpoint_restore(net, exp, "model", use_cuda=False) == 1      # nothing to restore -> epoch 1
    for epoch in (1, 2, 3, 4, 5):
        scn.checkpoint_save(net, exp, "model", epoch, use_cuda=False)
    files = sorted(os.listdir(tmp_path))
    # epoch-1 files survive only when epoch-1 is a power of two (train.py:91)
    assert files == ["exp-%09d-model.pth" % e for e in (1, 2, 4, 5)]
    other = build_encoder(scn, "SparseConvUNet", 16, 1, False)
    assert scn.checkpoint_restore(other, exp, "model", use_cuda=False) == 6
    for (k, a), (_, b) in zip(net.state_dict().items(), other.state_dict().items()):
        assert torch.equal(a, b), k
    assert scn.is_power2(8) and not scn.is_power2(0) and not scn.is_power2(6)


def test_op_counters_behave_like_plain_numbers():
    scn.forward_pass_multiplyAdd_count = 0
    scn.forward_pass_hidden_states = 0
    scn.forward_pass_multiplyAdd_count += 5
    scn._add_madds(7, 3)
    assert scn.forward_pass_multiplyAdd_count == 26
    assert scn.forward_pass_multiplyAdd_count / 2 / 1e6 == 13e-6      # train.py:86 style arithmetic
    scn.forward_pass_multiplyAdd_count = 0
    assert scn.forward_pass_multiplyAdd_count == 0


def test_gemm_weight_forms():
    w = torch.randn(27, 5, 7)
    x = torch.randn(4, 5)
    g = torch.randn(4, 7)
    f = GemmWeight(w)
    assert (f.cin, f.cout) == (5, 7)
    assert torch.allclose(x @ f.rowmajor()[3], x @ w[3])
    b = GemmWeight(w, transposed=True, flip=True)
    assert (b.cin, b.cout) == (7, 5)
    assert torch.allclose(g @ b.rowmajor()[3], g @ w[23].t())
    # the K-major tensor-core form is produced by a CUDA kernel: tests/test_gpu_ops.py::test_weight_preparation


def test_capacity_rounding_and_scene_generator():
    from sparseconvnet._lib import round_rows
    for n in (1, 255, 4096, 4097, 648937, 1200000):
        c = round_rows(n)
        assert n <= c <= n * 1.13 + 256
    assert round_rows(648937) == round_rows(650001)      # neighbouring step sizes share one buffer size
    xyz, rgb = make_scene(0, 5000)
    assert xyz.shape == (5000, 3) and abs(xyz.mean()) < 1e-9 and rgb.min() >= -1 and rgb.max() <= 1
    coords, feats, offs = make_batch([0, 1], 50, n_points=5000)
    assert coords.dtype == torch.int64 and coords.shape[1] == 4 and coords[:, :3].min() >= 0 and coords[:, :3].max() < 4096
    assert offs[-1] == coords.shape[0] == feats.shape[0]
    assert (coords[:offs[1], 3] == 0).all() and (coords[offs[1]:, 3] == 1).all()
    c2, _, _ = make_batch([0, 1], 50, n_points=5000, step=1)
    assert not torch.equal(coords[:100], c2[:100])       # a new step re-randomises the pose


def test_checkpoint_with_optimizer_and_scheduler_state(tmp_path):
    """f4: the reference saves only model.state_dict() (train.py:91) and restarts Adam from zero; the optional
    optimizer= / scheduler= arguments round-trip the training state too, pruned by the same power-of-two rule."""
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    sch = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.5)
    exp = os.path.join(tmp_path, "run")
    for epoch in (1, 2, 3):
        net(torch.randn(5, 4)).sum().backward()
        opt.step(); opt.zero_grad(); sch.step()
        scn.checkpoint_save(net, exp, "model", epoch, use_cuda=False, optimizer=opt, scheduler=sch)
    assert sorted(os.listdir(tmp_path)) == sorted(
        ["run-%09d-model.pth" % e for e in (1, 2, 3)] + ["run-%09d-model.train.pth" % e for e in (1, 2, 3)])
    scn.checkpoint_save(net, exp, "model", 4, use_cuda=False, optimizer=opt, scheduler=sch)   # epoch 3 is pruned, both files
    assert not os.path.exists(exp + "-%09d-model.pth" % 3) and not os.path.exists(exp + "-%09d-model.train.pth" % 3)
    net2 = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    opt2 = torch.optim.Adam(net2.parameters(), lr=1e-3)
    sch2 = torch.optim.lr_scheduler.StepLR(opt2, step_size=2, gamma=0.5)
    assert scn.checkpoint_restore(net2, exp, "model", use_cuda=False, optimizer=opt2, scheduler=sch2) == 5
    s1, s2 = opt.state_dict(), opt2.state_dict()
    assert s1["param_groups"] == s2["param_groups"]
    for k in s1["state"]:
        for name in ("exp_avg", "exp_avg_sq", "step"):
            assert torch.equal(torch.as_tensor(s1["state"][k][name]), torch.as_tensor(s2["state"][k][name]))
    assert sch2.state_dict() == sch.state_dict()
    # the upstream call (no optimizer) still works on the same directory and ignores the .train.pth files
    assert scn.checkpoint_restore(torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2)), exp, "model",
                                  use_cuda=False) == 5


def test_upstream_keyed_state_dict_loads():
    """f4: a state_dict with upstream sparseconvnet's key names and shapes (tests/golden/upstream_unet_keys.json:
    module-tree indices of scn.Sequential/UNet as Function_test.py:113-226 restates them, conv `weight`
    (filter_volume, 1, nIn, nOut), BatchNorm `weight, bias, runningMean, runningVar` in camelCase) loads
    strictly into this package's encoder and reaches the right tensors."""
    import json
    here = os.path.dirname(os.path.abspath(__file__))
    spec = json.load(open(os.path.join(here, "golden", "upstream_unet_keys.json")))
    net = build_encoder(scn, "SparseConvUNet", 16, 1, False)
    sd = {}
    for i, (k, shape) in enumerate(spec["keys"]):
        sd[k] = torch.full(shape, float(i % 97) + 1.0)
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    own = net.state_dict()
    assert len(own) == len(sd)
    for i, (k, shape) in enumerate(spec["keys"]):
        k2 = k.replace("runningMean", "running_mean").replace("runningVar", "running_var")
        assert tuple(own[k2].shape) == tuple(shape) and float(own[k2].flatten()[0]) == float(i % 97) + 1.0, k

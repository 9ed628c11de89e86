"""f1: fused scene head (b200scn_heads.py, csrc/head.cu) against the reference's own composition -- per-point OutputLayer,
the Python loop of torch.mean per scene (models/SparseConvNet.py:20-26), nn.Linear (models/MultiLabelContrastive.py:66)
and F.multilabel_soft_margin_loss (utils/loss.py:28) -- run on the CPU oracle."""
import pytest
import torch
import torch.nn.functional as F

from _util import copy_params, random_cloud, rel_err

pytestmark = pytest.mark.gpu


def test_head_kernels_match_torch():
    from b200scn_heads import MultiLabelHeadFn
    torch.manual_seed(0)
    for B, C, NC, bias in ((8, 256, 20, True), (5, 448, 20, True), (1, 32, 20, False), (30, 128, 7, True)):
        pooled = torch.randn(B, C, device="cuda", requires_grad=True)
        w = (torch.randn(NC, C, device="cuda") * 0.1).requires_grad_(True)
        b = torch.randn(NC, device="cuda").requires_grad_(True) if bias else None
        labels = (torch.rand(B, NC, device="cuda") < 0.3).float()
        logits, loss = MultiLabelHeadFn.apply(pooled, w, b, labels)
        pr = pooled.detach().double().requires_grad_(True)
        wr = w.detach().double().requires_grad_(True)
        br = b.detach().double().requires_grad_(True) if bias else None
        lr = F.linear(pr, wr, br)
        ref_loss = F.multilabel_soft_margin_loss(lr, labels.double())
        assert rel_err(logits, lr) < 1e-5
        assert abs(float(loss) - float(ref_loss)) < 1e-5 * max(1.0, abs(float(ref_loss)))
        # the loss and an extra use of the logits both feed the backward
        gl = torch.randn(B, NC, device="cuda")
        (loss * 3.0 + (logits * gl).sum()).backward()
        (ref_loss * 3.0 + (lr * gl.double()).sum()).backward()
        assert rel_err(pooled.grad, pr.grad) < 1e-5
        assert rel_err(w.grad, wr.grad) < 1e-5
        if bias:
            assert rel_err(b.grad, br.grad) < 1e-5
        # bit-reproducible loss
        _, loss2 = MultiLabelHeadFn.apply(pooled.detach(), w.detach(), b.detach() if bias else None, labels)
        assert torch.equal(loss2, loss.detach())
        # logits only (no labels): loss is not computed, gradient = d_logits only
        p2 = pooled.detach().clone().requires_grad_(True)
        lg, _ = MultiLabelHeadFn.apply(p2, w.detach(), b.detach() if bias else None, None)
        (lg * gl).sum().backward()
        assert rel_err(p2.grad, gl.double() @ w.detach().double()) < 1e-5


@pytest.mark.parametrize("kind,m,res", [("SparseConvUNet", 8, False), ("SparseConvFCNetDirectUpPool", 8, True)])
def test_multilabel_model_matches_reference_composition(kind, m, res):
    import sparseconvnet as scn
    from b200scn_heads import MultiLabelHead
    from b200scn_synth import EMBED_WIDTH, build_encoder
    from oracle import scn_oracle as ref
    scn.set_precision("fp32")
    torch.manual_seed(1)
    coords, feats = random_cloud(5, 1500, 24, 3, dup_frac=0.4)
    offs = [0, 1500, 3000, 4500]
    enc_r = build_encoder(ref, kind, m, 1, res, full_scale=4096)
    enc_g = build_encoder(scn, kind, m, 1, res, full_scale=4096)
    for net in (enc_r, enc_g):
        for mod in net.modules():
            if hasattr(mod, "leakiness"):
                mod.leakiness = 1.0      # continuous gradients (mask flips: tests/test_gpu_nets.py)
    copy_params(enc_r, enc_g)
    width = EMBED_WIDTH[kind](m)
    lin_r = torch.nn.Linear(width, 20)
    model = MultiLabelHead(enc_g, width)
    model.linear.load_state_dict(lin_r.state_dict())
    model.cuda()
    assert set(k.split(".")[0] for k in model.state_dict()) == {"pc_encoder", "linear"}   # the reference's key prefixes
    labels = (torch.rand(3, 20) < 0.3).float()
    # reference composition on the CPU oracle
    fr = feats.clone().requires_grad_(True)
    out = enc_r([coords, fr])
    g_feats = torch.stack([out[offs[i]:offs[i + 1]].mean(0) for i in range(3)])
    logits_r = lin_r(g_feats)
    loss_r = F.multilabel_soft_margin_loss(logits_r, labels)
    loss_r.backward()
    # fused path
    fg = feats.cuda().requires_grad_(True)
    batch = {"coords": coords, "feature": fg, "batch_offsets": offs}
    logits_g, loss_g = model((batch, None), istrain=True, labels=labels.cuda())
    loss_g.backward()
    assert rel_err(logits_g, logits_r) < 1e-4
    assert abs(float(loss_g) - float(loss_r)) < 1e-5
    assert rel_err(fg.grad, fr.grad) < 1e-3
    assert rel_err(model.linear.weight.grad, lin_r.weight.grad) < 1e-3
    assert rel_err(model.linear.bias.grad, lin_r.bias.grad) < 1e-3
    for (k, pg), (_, pr) in zip(enc_g.named_parameters(), enc_r.named_parameters()):
        assert rel_err(pg.grad, pr.grad) < 2e-3, k
    # without labels the training forward returns (logits, None) exactly like the reference's MultiLabel.forward
    lg, none = model((batch, None), istrain=True)
    assert none is None and rel_err(lg, logits_r) < 1e-4
    # evaluation: per-point logits (models/MultiLabelContrastive.py:62-70 with istrain=False)
    enc_r([coords, feats.clone()])      # second training forward, as the fused model has done: same running statistics
    model.eval(); enc_r.eval()
    with torch.no_grad():
        per_point = model(batch)
        ref_pp = lin_r(enc_r([coords, feats]))
    assert per_point.shape == (4500, 20) and rel_err(per_point, ref_pp) < 1e-4

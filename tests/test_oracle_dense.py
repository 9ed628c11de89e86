"""Pin the oracle's floating-point half against independent dense implementations (torch conv3d /
conv_transpose3d / manual batch norm) and fp64 gradcheck -- SURVEY 8c golden-vector plan (ii)."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import scn_oracle as ref


def _full_grid(S, B, C, seed=0, dtype=torch.float64):
    g = torch.stack(torch.meshgrid(torch.arange(S), torch.arange(S), torch.arange(S), indexing="ij"), -1).reshape(-1, 3)
    coords = torch.cat([torch.cat([g, torch.full((g.shape[0], 1), b)], 1) for b in range(B)])
    torch.manual_seed(seed)
    coords = coords[torch.randperm(coords.shape[0])]
    return coords, torch.randn(coords.shape[0], C, dtype=dtype)


def test_subm_equals_conv3d_on_full_grid():
    coords, feats = _full_grid(6, 2, 5)
    x = ref.InputLayer(3, 6, mode=4)([coords, feats])
    conv = ref.SubmanifoldConvolution(3, 5, 7, 3, False).double()
    dense = ref.SparseToDense(3, 7)(conv(x))
    w = conv.weight.view(3, 3, 3, 5, 7).permute(4, 3, 0, 1, 2)
    assert torch.allclose(dense, F.conv3d(ref.SparseToDense(3, 5)(x), w, padding=1), atol=1e-12)


def test_subm_on_sparse_set_equals_masked_conv3d():
    torch.manual_seed(3)
    S = 8
    coords, feats = _full_grid(S, 1, 4)
    keep = torch.rand(coords.shape[0]) < 0.35
    coords, feats = coords[keep], feats[keep]
    x = ref.InputLayer(3, S, mode=4)([coords, feats])
    conv = ref.SubmanifoldConvolution(3, 4, 6, 3, False).double()
    y = conv(x)
    w = conv.weight.view(3, 3, 3, 4, 6).permute(4, 3, 0, 1, 2)
    full = F.conv3d(ref.SparseToDense(3, 4)(x), w, padding=1)
    vox = torch.from_numpy(x.metadata.vox[S].astype(np.int64))
    assert torch.allclose(y.features, full[0][:, vox[:, 0], vox[:, 1], vox[:, 2]].t(), atol=1e-12)


def test_strided_conv_and_deconv_equal_dense():
    for s in (2, 4):
        coords, feats = _full_grid(8, 2, 3)
        x = ref.InputLayer(3, 8, mode=4)([coords, feats])
        c = ref.Convolution(3, 3, 4, s, s, False).double()
        y = c(x)
        d = ref.SparseToDense(3, 4)(y)
        w = c.weight.view(s, s, s, 3, 4).permute(4, 3, 0, 1, 2)
        assert torch.allclose(d, F.conv3d(ref.SparseToDense(3, 3)(x), w, stride=s), atol=1e-12)
        dc = ref.Deconvolution(3, 4, 5, s, s, False).double()
        z = dc(y)
        assert int(z.spatial_size[0]) == 8
        w2 = dc.weight.view(s, s, s, 4, 5).permute(3, 4, 0, 1, 2)
        assert torch.allclose(ref.SparseToDense(3, 5)(z), F.conv_transpose3d(d, w2, stride=s), atol=1e-12)
        up = ref.UnPooling(3, s, s)(y)
        assert torch.allclose(ref.SparseToDense(3, 4)(up), d.repeat_interleave(s, 2).repeat_interleave(s, 3).repeat_interleave(s, 4))


def test_batchnorm_matches_manual_formula_and_torch():
    torch.manual_seed(0)
    x = torch.randn(50, 6, dtype=torch.float64) * 3 + 1
    t = ref.SparseConvNetTensor(x.clone().requires_grad_(True), None, torch.LongTensor([8] * 3))
    bn = ref.BatchNormLeakyReLU(6, leakiness=0.2).double()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 2)
        bn.bias.uniform_(-1, 1)
    y = bn(t).features
    x2 = x.clone().requires_grad_(True)
    y2 = F.leaky_relu(F.batch_norm(x2, None, None, bn.weight, bn.bias, True, 0.0, 1e-4), 0.2)
    assert torch.allclose(y, y2, atol=1e-10)
    go = torch.randn_like(y)
    y.backward(go)
    y2.backward(go)
    assert torch.allclose(t.features.grad, x2.grad, atol=1e-9)
    # running statistics: 0.9*old + 0.1*batch, unbiased variance (App. B.8)
    assert torch.allclose(bn.running_mean, 0.1 * x.mean(0))
    assert torch.allclose(bn.running_var, 0.9 + 0.1 * x.var(0, unbiased=True))


def test_explicit_backward_matches_autograd_fp64():
    torch.manual_seed(1)
    coords = torch.cat([torch.randint(0, 6, (60, 3)), torch.randint(0, 2, (60, 1))], 1)
    feats = torch.randn(60, 3, dtype=torch.float64, requires_grad=True)
    net = ref.Sequential(
        ref.InputLayer(3, 8, mode=4), ref.SubmanifoldConvolution(3, 3, 4, 3, False),
        ref.UNet(3, 1, [4, 6], residual_blocks=True), ref.BatchNormReLU(4), ref.OutputLayer(3)).double()
    out = net([coords, feats])
    (out ** 2).sum().backward()
    g_explicit = [p.grad.clone() for p in net.parameters()] + [feats.grad.clone()]
    # finite differences on a few weights
    params = list(net.parameters())
    eps = 1e-6
    for pi in (0, 3, len(params) - 3):
        p = params[pi]
        flat = p.data.view(-1)
        for j in (0, flat.numel() // 2):
            old = flat[j].item()
            flat[j] = old + eps
            lp = (net([coords, feats.detach()]) ** 2).sum().item()
            flat[j] = old - eps
            lm = (net([coords, feats.detach()]) ** 2).sum().item()
            flat[j] = old
            fd = (lp - lm) / (2 * eps)
            assert abs(fd - g_explicit[pi].view(-1)[j].item()) < 1e-4 * max(1.0, abs(fd)), (pi, j)

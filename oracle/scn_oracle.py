"""oracle/scn_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle; never imported by the product).

CPU restatement of the sparseconvnet 0.2 ("scn") semantics behind the reference's
SparseConvUNet / SparseConvFCNet* encoders (models/SparseConvNet.py:57-211).  The real scn is
an absent third-party dependency (requirements.txt:2) and the reference pins none of its
results in tests, so this oracle is PARITY UNPINNED against upstream; it follows
SURVEY.md App. B and is pinned by hand-written literal cases, dense conv3d / conv_transpose3d
equivalence and fp64 gradcheck (tests/test_oracle_*.py).

Algorithm shape = upstream's CPU path: hash-map rulebooks (oracle/scn_rules.c), then per
kernel offset `index_select -> matmul -> index_add_`, explicit backward formulas
(App. B.4, B.8), torch CPU tensors (MKL/OpenMP threads).  It exposes the same module surface
as scn so models/SparseConvNet.py composes it unchanged, and doubles as the timed CPU baseline
(bench.py `cpu_baseline` / `--impl reference`).
"""
import ctypes
import glob
import os
import sys

import numpy as np
import torch

from . import build as _build

_lib = ctypes.CDLL(_build.build())
_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)
_f32p = ctypes.POINTER(ctypes.c_float)
_lib.oracle_input_rules.restype = ctypes.c_int64
_lib.oracle_input_rules.argtypes = [_i64p, ctypes.c_int64, ctypes.c_int, _i32p, _i32p]
_lib.oracle_subm_map.restype = ctypes.c_int
_lib.oracle_subm_map.argtypes = [_i32p, ctypes.c_int64, _i32p]
_lib.oracle_strided.restype = ctypes.c_int64
_lib.oracle_strided.argtypes = [_i32p, ctypes.c_int64, ctypes.c_int, _i32p, _i32p, _i32p]
_lib.oracle_ball_query.restype = None
_lib.oracle_ball_query.argtypes = [ctypes.c_int] * 3 + [ctypes.c_float, ctypes.c_int, _f32p, _f32p, _i32p, _i32p]
_lib.oracle_group_points.restype = None
_lib.oracle_group_points.argtypes = [ctypes.c_int] * 5 + [_f32p, _i32p, _f32p]
_lib.oracle_group_points_grad.restype = None
_lib.oracle_group_points_grad.argtypes = [ctypes.c_int] * 5 + [_f32p, _i32p, _f32p]

# op counters, same meaning as scn's module globals (train.py:50-51,86-87)
forward_pass_multiplyAdd_count = 0
forward_pass_hidden_states = 0
_this = sys.modules[__name__]


def _p(a, t):
    return a.ctypes.data_as(t)


# ----------------------------------------------------------------------------- rulebooks
def input_rules(coords: np.ndarray):
    """coords (P,3|4) int64 -> (pv (P,) int32, vox (N,4) int32 [x,y,z,b])."""
    coords = np.ascontiguousarray(coords, dtype=np.int64)
    P, ncols = coords.shape
    pv = np.empty(P, np.int32)
    vox = np.empty((max(P, 1), 4), np.int32)
    n = _lib.oracle_input_rules(_p(coords, _i64p), P, ncols, _p(pv, _i32p), _p(vox, _i32p))
    if n < 0:
        raise ValueError("oracle_input_rules: coordinates out of range")
    return pv, vox[:n].copy()


def subm_map(vox: np.ndarray):
    """(N,4) sites -> nbr (N,27) int32, -1 = absent (App. B.5)."""
    vox = np.ascontiguousarray(vox, dtype=np.int32)
    nbr = np.empty((vox.shape[0], 27), np.int32)
    if vox.shape[0]:
        _lib.oracle_subm_map(_p(vox, _i32p), vox.shape[0], _p(nbr, _i32p))
    return nbr


def strided(vox: np.ndarray, s: int):
    """(Nf,4) fine sites -> parent (Nf,), off (Nf,), coarse vox (Nc,4)  (App. B.6)."""
    vox = np.ascontiguousarray(vox, dtype=np.int32)
    nf = vox.shape[0]
    parent = np.empty(nf, np.int32)
    off = np.empty(nf, np.int32)
    voxc = np.empty((max(nf, 1), 4), np.int32)
    nc = _lib.oracle_strided(_p(vox, _i32p), nf, s, _p(parent, _i32p), _p(off, _i32p), _p(voxc, _i32p))
    return parent, off, voxc[:nc].copy()


def rules_from_map(nbr: np.ndarray):
    """Neighbour table -> scn-style rulebook: list over offsets of (in_ids, out_ids), ascending out."""
    rules = []
    for k in range(nbr.shape[1]):
        out = np.nonzero(nbr[:, k] >= 0)[0].astype(np.int64)
        rules.append((nbr[out, k].astype(np.int64), out))
    return rules


def rules_from_parent(parent: np.ndarray, off: np.ndarray, K: int):
    """Strided rulebook: per offset (fine ids, coarse ids), ascending fine id."""
    rules = []
    for k in range(K):
        f = np.nonzero(off == k)[0].astype(np.int64)
        rules.append((f, parent[f].astype(np.int64)))
    return rules


class Metadata:
    """Per-forward cache of grids and rulebooks, keyed by spatial size (App. B.1)."""

    def __init__(self, dimension=3):
        self.dimension = dimension
        self.vox = {}        # spatial size -> (N,4) int32
        self.subm = {}       # spatial size -> list of (in,out) LongTensors
        self.subm_nbr = {}   # spatial size -> (N,27) int32 numpy
        self.down = {}       # (fine size, s) -> dict(parent, off, rules)
        self.pv = None       # point -> voxel (P,) int64 tensor
        self.counts = None   # (N0,) points per voxel
        self.mode = 4

    def get_subm_rules(self, size):
        if size not in self.subm:
            nbr = subm_map(self.vox[size])
            self.subm_nbr[size] = nbr
            self.subm[size] = [(torch.from_numpy(i), torch.from_numpy(o)) for i, o in rules_from_map(nbr)]
        return self.subm[size]

    def get_down_rules(self, size, s):
        if (size, s) not in self.down:
            assert size % s == 0, "spatial size must be divisible by the stride"
            parent, off, voxc = strided(self.vox[size], s)
            csize = (size - s) // s + 1
            self.vox.setdefault(csize, voxc)
            rules = [(torch.from_numpy(f), torch.from_numpy(c)) for f, c in rules_from_parent(parent, off, s ** 3)]
            self.down[(size, s)] = dict(parent=parent, off=off, rules=rules, nc=voxc.shape[0], csize=csize)
        return self.down[(size, s)]


class SparseConvNetTensor:
    def __init__(self, features=None, metadata=None, spatial_size=None):
        self.features = features
        self.metadata = metadata
        self.spatial_size = spatial_size

    def size(self):
        return int(self.spatial_size[0])


# ----------------------------------------------------------------------------- autograd functions
class _InputFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, pv, mult, n):
        out = torch.zeros(n, feats.shape[1], dtype=feats.dtype)
        out.index_add_(0, pv, feats * mult[:, None])
        ctx.save_for_backward(pv, mult)
        return out

    @staticmethod
    def backward(ctx, g):
        pv, mult = ctx.saved_tensors
        return g.index_select(0, pv) * mult[:, None], None, None, None


class _OutputFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, pv, sel):
        ctx.save_for_backward(pv, sel)
        ctx.n = feats.shape[0]
        return feats.index_select(0, pv) * sel[:, None]

    @staticmethod
    def backward(ctx, g):
        pv, sel = ctx.saved_tensors
        d = torch.zeros(ctx.n, g.shape[1], dtype=g.dtype)
        d.index_add_(0, pv, g * sel[:, None])
        return d, None, None


class _RuleConvFn(torch.autograd.Function):
    """out[o] += in[i] @ W[k] over the pairs (i,o) of rules[k] (App. B.4 / B.6 / B.7)."""

    @staticmethod
    def forward(ctx, x, w, rules, n_out):
        out = torch.zeros(n_out, w.shape[-1], dtype=x.dtype)
        for k, (ri, ro) in enumerate(rules):
            if ri.numel():
                out.index_add_(0, ro, x.index_select(0, ri) @ w[k])
        ctx.rules = rules
        ctx.save_for_backward(x, w)
        return out

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        dx = torch.zeros_like(x)
        dw = torch.zeros_like(w)
        for k, (ri, ro) in enumerate(ctx.rules):
            if ri.numel():
                gk = g.index_select(0, ro)
                dx.index_add_(0, ri, gk @ w[k].t())
                dw[k] = x.index_select(0, ri).t() @ gk
        return dx, dw, None, None


class _UnPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, parent, n_fine):
        ctx.save_for_backward(parent)
        ctx.nc = x.shape[0]
        return x.index_select(0, parent)

    @staticmethod
    def backward(ctx, g):
        (parent,) = ctx.saved_tensors
        d = torch.zeros(ctx.nc, g.shape[1], dtype=g.dtype)
        d.index_add_(0, parent, g)
        return d, None, None


class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, parent, nc):
        out = torch.zeros(nc, x.shape[1], dtype=x.dtype)
        idx = parent[:, None].expand(-1, x.shape[1])
        out = out.scatter_reduce(0, idx, x, reduce="amax", include_self=True)
        ctx.save_for_backward(x, out, parent)
        return out

    @staticmethod
    def backward(ctx, g):
        x, out, parent = ctx.saved_tensors
        return g.index_select(0, parent) * (x == out.index_select(0, parent)).to(g.dtype), None, None


class _BNFn(torch.autograd.Function):
    """App. B.8 (eps 1e-4, momentum 0.9 on the OLD value, unbiased running var, leaky ReLU)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, train, leak):
        n = x.shape[0]
        if train:
            mean = x.sum(0) / n
            var = (x * x).sum(0) / n - mean * mean
            var = torch.clamp(var, min=0)
            running_mean.mul_(momentum).add_((1 - momentum) * mean.detach())
            running_var.mul_(momentum).add_((1 - momentum) * var.detach() * (n / max(n - 1, 1)))
            invstd = (var + eps).rsqrt()
        else:
            mean = running_mean.clone()
            invstd = (running_var + eps).rsqrt()
        y = (x - mean) * invstd * weight + bias
        out = torch.where(y > 0, y, leak * y)
        ctx.save_for_backward(x, out, weight, mean, invstd)
        ctx.leak, ctx.train = leak, bool(train)
        return out

    @staticmethod
    def backward(ctx, go):
        x, out, weight, mean, invstd = ctx.saved_tensors
        n = x.shape[0]
        g = go * torch.where(out > 0, torch.ones_like(out), torch.full_like(out, ctx.leak))
        d_bias = g.sum(0)
        xc = x - mean
        dotp = (xc * g).sum(0)
        d_weight = dotp * invstd
        if ctx.train:
            d_in = (g - d_bias / n - xc * dotp * invstd * invstd / n) * invstd * weight
        else:   # eval: mean / invstd are the running statistics, constants with respect to x (upstream asserts train here)
            d_in = g * invstd * weight
        return d_in, d_weight, d_bias, None, None, None, None, None, None


# ----------------------------------------------------------------------------- modules
class Sequential(torch.nn.Sequential):
    def add(self, module):
        self._modules[str(len(self._modules))] = module
        return self

    def forward(self, input):
        for m in self._modules.values():
            input = m(input)
        return input


class Identity(torch.nn.Module):
    def forward(self, input):
        return input


class ConcatTable(Sequential):
    def forward(self, input):
        return [m(input) for m in self._modules.values()]


class AddTable(torch.nn.Module):
    def forward(self, input):
        out = SparseConvNetTensor(None, input[0].metadata, input[0].spatial_size)
        out.features = sum(i.features for i in input)
        return out


class JoinTable(torch.nn.Module):
    def forward(self, input):
        out = SparseConvNetTensor(None, input[0].metadata, input[0].spatial_size)
        out.features = torch.cat([i.features for i in input], 1)
        return out


class InputLayer(torch.nn.Module):
    def __init__(self, dimension, spatial_size, mode=3):
        super().__init__()
        self.dimension = dimension
        self.spatial_size = torch.LongTensor([int(spatial_size)] * dimension) if np.isscalar(spatial_size) else torch.LongTensor(list(spatial_size))
        self.mode = mode

    def forward(self, input):
        coords, feats = input[0], input[1]
        md = Metadata(self.dimension)
        c = coords.cpu().long().numpy()
        assert (c[:, :3] < int(self.spatial_size[0])).all(), "coordinate outside spatial_size"
        pv, vox = input_rules(c)
        size = int(self.spatial_size[0])
        md.vox[size] = vox
        n = vox.shape[0]
        pvt = torch.from_numpy(pv.astype(np.int64))
        counts = torch.bincount(pvt, minlength=n)
        P = pvt.numel()
        rows = torch.arange(P)
        first = torch.full((n,), P, dtype=torch.long).scatter_reduce(0, pvt, rows, "amin")
        last = torch.full((n,), -1, dtype=torch.long).scatter_reduce(0, pvt, rows, "amax")
        dt = feats.dtype
        if self.mode == 4:
            mult = (1.0 / counts.to(dt))[pvt]
        elif self.mode == 3:
            mult = torch.ones(P, dtype=dt)
        elif self.mode == 2:
            mult = (first[pvt] == rows).to(dt)
        elif self.mode == 1:
            mult = (last[pvt] == rows).to(dt)
        else:
            raise NotImplementedError("InputLayer mode 0")
        md.pv, md.counts, md.mode = pvt, counts, self.mode
        # OutputLayer row selector: all rows for modes 3/4, only the kept row for 1/2 (App. B.3)
        md.out_sel = torch.ones(P, dtype=dt) if self.mode in (3, 4) else mult.clone()
        out = SparseConvNetTensor(None, md, self.spatial_size)
        out.features = _InputFn.apply(feats.cpu(), pvt, mult, n)
        return out


class OutputLayer(torch.nn.Module):
    def __init__(self, dimension):
        super().__init__()
        self.dimension = dimension

    def forward(self, input):
        md = input.metadata
        return _OutputFn.apply(input.features, md.pv, md.out_sel.to(input.features.dtype))


class SubmanifoldConvolution(torch.nn.Module):
    def __init__(self, dimension, nIn, nOut, filter_size, bias, groups=1):
        super().__init__()
        assert groups == 1 and filter_size == 3 and dimension == 3
        self.nIn, self.nOut = nIn, nOut
        self.filter_volume = 27
        std = (2.0 / nIn / self.filter_volume) ** 0.5
        self.weight = torch.nn.Parameter(torch.Tensor(27, 1, nIn, nOut).normal_(0, std))
        self.bias = torch.nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, input):
        assert input.features.shape[1] == self.nIn
        rules = input.metadata.get_subm_rules(input.size())
        out = SparseConvNetTensor(None, input.metadata, input.spatial_size)
        w = self.weight.view(27, self.nIn, self.nOut)
        out.features = _RuleConvFn.apply(input.features, w, rules, input.features.shape[0])
        if self.bias is not None:
            out.features = out.features + self.bias
        _this.forward_pass_multiplyAdd_count += sum(r[0].numel() for r in rules) * self.nIn * self.nOut
        _this.forward_pass_hidden_states += out.features.nelement()
        return out


class Convolution(torch.nn.Module):
    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        super().__init__()
        assert groups == 1 and filter_size == filter_stride and dimension == 3
        self.nIn, self.nOut, self.s = nIn, nOut, filter_size
        K = filter_size ** 3
        std = (2.0 / nIn / K) ** 0.5
        self.weight = torch.nn.Parameter(torch.Tensor(K, 1, nIn, nOut).normal_(0, std))
        self.bias = torch.nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, input):
        assert input.features.shape[1] == self.nIn
        d = input.metadata.get_down_rules(input.size(), self.s)
        out = SparseConvNetTensor(None, input.metadata, torch.LongTensor([d["csize"]] * 3))
        w = self.weight.view(-1, self.nIn, self.nOut)
        out.features = _RuleConvFn.apply(input.features, w, d["rules"], d["nc"])
        if self.bias is not None:
            out.features = out.features + self.bias
        _this.forward_pass_multiplyAdd_count += input.features.shape[0] * self.nIn * self.nOut
        _this.forward_pass_hidden_states += out.features.nelement()
        return out


class Deconvolution(torch.nn.Module):
    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        super().__init__()
        assert groups == 1 and filter_size == filter_stride and dimension == 3
        self.nIn, self.nOut, self.s = nIn, nOut, filter_size
        K = filter_size ** 3
        std = (2.0 / nIn / K) ** 0.5
        self.weight = torch.nn.Parameter(torch.Tensor(K, 1, nIn, nOut).normal_(0, std))
        self.bias = torch.nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, input):
        assert input.features.shape[1] == self.nIn
        fsize = (input.size() - 1) * self.s + self.s
        d = input.metadata.get_down_rules(fsize, self.s)
        out = SparseConvNetTensor(None, input.metadata, torch.LongTensor([fsize] * 3))
        w = self.weight.view(-1, self.nIn, self.nOut)
        rev = [(c, f) for f, c in d["rules"]]  # roles swapped: gather coarse, scatter fine (App. B.7)
        nf = input.metadata.vox[fsize].shape[0]
        out.features = _RuleConvFn.apply(input.features, w, rev, nf)
        if self.bias is not None:
            out.features = out.features + self.bias
        _this.forward_pass_multiplyAdd_count += nf * self.nIn * self.nOut
        _this.forward_pass_hidden_states += out.features.nelement()
        return out


class UnPooling(torch.nn.Module):
    def __init__(self, dimension, pool_size, pool_stride):
        super().__init__()
        assert pool_size == pool_stride
        self.s = pool_size

    def forward(self, input):
        fsize = (input.size() - 1) * self.s + self.s
        d = input.metadata.get_down_rules(fsize, self.s)
        out = SparseConvNetTensor(None, input.metadata, torch.LongTensor([fsize] * 3))
        parent = torch.from_numpy(d["parent"].astype(np.int64))
        out.features = _UnPoolFn.apply(input.features, parent, parent.numel())
        return out


class MaxPooling(torch.nn.Module):
    def __init__(self, dimension, pool_size, pool_stride):
        super().__init__()
        assert pool_size == pool_stride
        self.s = pool_size

    def forward(self, input):
        d = input.metadata.get_down_rules(input.size(), self.s)
        out = SparseConvNetTensor(None, input.metadata, torch.LongTensor([d["csize"]] * 3))
        parent = torch.from_numpy(d["parent"].astype(np.int64))
        out.features = _MaxPoolFn.apply(input.features, parent, d["nc"])
        return out


class BatchNormalization(torch.nn.Module):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, affine=True, leakiness=1):
        super().__init__()
        self.nPlanes, self.eps, self.momentum, self.leakiness = nPlanes, eps, momentum, leakiness
        self.register_buffer("running_mean", torch.zeros(nPlanes))
        self.register_buffer("running_var", torch.ones(nPlanes))
        self.weight = torch.nn.Parameter(torch.ones(nPlanes))
        self.bias = torch.nn.Parameter(torch.zeros(nPlanes))

    def forward(self, input):
        assert input.features.shape[1] == self.nPlanes
        out = SparseConvNetTensor(None, input.metadata, input.spatial_size)
        out.features = _BNFn.apply(input.features, self.weight, self.bias, self.running_mean, self.running_var,
                                   self.eps, self.momentum, self.training, self.leakiness)
        return out


class BatchNormReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9):
        super().__init__(nPlanes, eps, momentum, True, 0)


class BatchNormLeakyReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, leakiness=0.333):
        super().__init__(nPlanes, eps, momentum, True, leakiness)


class NetworkInNetwork(torch.nn.Module):
    def __init__(self, nIn, nOut, bias):
        super().__init__()
        self.nIn, self.nOut = nIn, nOut
        self.weight = torch.nn.Parameter(torch.Tensor(nIn, nOut).normal_(0, (2.0 / nIn) ** 0.5))
        self.bias = torch.nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, input):
        assert input.features.shape[1] == self.nIn
        out = SparseConvNetTensor(None, input.metadata, input.spatial_size)
        out.features = input.features @ self.weight
        if self.bias is not None:
            out.features = out.features + self.bias
        _this.forward_pass_multiplyAdd_count += input.features.shape[0] * self.nIn * self.nOut
        _this.forward_pass_hidden_states += out.features.nelement()
        return out


class SparseToDense(torch.nn.Module):
    def __init__(self, dimension, nPlanes):
        super().__init__()
        self.nPlanes = nPlanes

    def forward(self, input):
        vox = torch.from_numpy(input.metadata.vox[input.size()].astype(np.int64))
        s = input.size()
        B = int(vox[:, 3].max()) + 1 if vox.numel() else 0
        dense = torch.zeros(B, s, s, s, self.nPlanes, dtype=input.features.dtype)
        dense = dense.index_put((vox[:, 3], vox[:, 0], vox[:, 1], vox[:, 2]), input.features)
        return dense.permute(0, 4, 1, 2, 3).contiguous()


# ----------------------------------------------------------------------------- net builders
def _block_factory(ns, dimension, residual_blocks, bn):
    def block(m, a, b):
        if residual_blocks:
            m.add(ns.ConcatTable()
                  .add(ns.Identity() if a == b else ns.NetworkInNetwork(a, b, False))
                  .add(ns.Sequential()
                       .add(bn(a)).add(ns.SubmanifoldConvolution(dimension, a, b, 3, False))
                       .add(bn(b)).add(ns.SubmanifoldConvolution(dimension, b, b, 3, False)))
                  ).add(ns.AddTable())
        else:
            m.add(ns.Sequential().add(bn(a)).add(ns.SubmanifoldConvolution(dimension, a, b, 3, False)))
    return block


def UNet(dimension, reps, nPlanes, residual_blocks=False, downsample=[2, 2], leakiness=0, n_input_planes=-1):
    """scn.UNet as called at models/SparseConvNet.py:63-68 (encoder half mirrored at Function_test.py:113-164;
    decoder half = BN + Deconvolution + JoinTable + blocks)."""
    ns = _this
    bn = lambda c: ns.BatchNormLeakyReLU(c, leakiness=leakiness)
    block = _block_factory(ns, dimension, residual_blocks, bn)

    def U(nPlanes, n_input_planes=-1):
        m = ns.Sequential()
        for i in range(reps):
            block(m, n_input_planes if n_input_planes != -1 else nPlanes[0], nPlanes[0])
            n_input_planes = -1
        if len(nPlanes) > 1:
            m.add(ns.ConcatTable().add(ns.Identity()).add(
                ns.Sequential()
                .add(bn(nPlanes[0]))
                .add(ns.Convolution(dimension, nPlanes[0], nPlanes[1], downsample[0], downsample[1], False))
                .add(U(nPlanes[1:]))
                .add(bn(nPlanes[1]))
                .add(ns.Deconvolution(dimension, nPlanes[1], nPlanes[0], downsample[0], downsample[1], False))))
            m.add(ns.JoinTable())
            for i in range(reps):
                block(m, nPlanes[0] * (2 if i == 0 else 1), nPlanes[0])
        return m
    return U(nPlanes, n_input_planes)


def FullyConvolutionalNet(dimension, reps, nPlanes, residual_blocks=False, downsample=[2, 2]):
    """scn.FullyConvolutionalNet as called at models/SparseConvNet.py:79-85 (variant mirrored at Function_test.py:166-226)."""
    ns = _this
    block = _block_factory(ns, dimension, residual_blocks, lambda c: ns.BatchNormReLU(c))

    def U(nPlanes):
        m = ns.Sequential()
        for _ in range(reps):
            block(m, nPlanes[0], nPlanes[0])
        if len(nPlanes) > 1:
            m.add(ns.ConcatTable().add(ns.Identity()).add(
                ns.Sequential()
                .add(ns.BatchNormReLU(nPlanes[0]))
                .add(ns.Convolution(dimension, nPlanes[0], nPlanes[1], downsample[0], downsample[1], False))
                .add(U(nPlanes[1:]))
                .add(ns.UnPooling(dimension, downsample[0], downsample[1]))))
            m.add(ns.JoinTable())
        return m
    return U(nPlanes)


# ----------------------------------------------------------------------------- utils (App. B.11)
def is_power2(num):
    return num != 0 and ((num & (num - 1)) == 0)


def checkpoint_restore(model, exp_name, name2, use_cuda=True, epoch=0):
    if epoch > 0:
        f = exp_name + "-%09d-" % epoch + name2 + ".pth"
        model.load_state_dict(torch.load(f))
    else:
        f = sorted(glob.glob(exp_name + "-*-" + name2 + ".pth"))
        if len(f) > 0:
            f = f[-1]
            model.load_state_dict(torch.load(f))
            epoch = int(f[len(exp_name) + 1:-len(name2) - 5])
    return epoch + 1


def checkpoint_save(model, exp_name, name2, epoch, use_cuda=True):
    f = exp_name + "-%09d-" % epoch + name2 + ".pth"
    torch.save(model.state_dict(), f)
    epoch = epoch - 1
    f = exp_name + "-%09d-" % epoch + name2 + ".pth"
    if os.path.isfile(f) and not is_power2(epoch):
        os.remove(f)


# ----------------------------------------------------------------------------- point2mask (A12)
def ball_query(radius, nsample, xyz, new_xyz, pointnums):
    xyz = np.ascontiguousarray(xyz, np.float32); new_xyz = np.ascontiguousarray(new_xyz, np.float32)
    pointnums = np.ascontiguousarray(pointnums, np.int32)
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = np.full((b, m, nsample), -1, np.int32)
    _lib.oracle_ball_query(b, n, m, float(radius), nsample, _p(new_xyz, _f32p), _p(xyz, _f32p), _p(pointnums, _i32p), _p(idx, _i32p))
    return idx


def group_points(points, idx):
    points = np.ascontiguousarray(points, np.float32); idx = np.ascontiguousarray(idx, np.int32)
    b, c, n = points.shape
    _, npoints, nsample = idx.shape
    out = np.zeros((b, c, npoints, nsample), np.float32)
    _lib.oracle_group_points(b, c, n, npoints, nsample, _p(points, _f32p), _p(idx, _i32p), _p(out, _f32p))
    return out


def group_points_grad(grad_out, idx, n):
    grad_out = np.ascontiguousarray(grad_out, np.float32); idx = np.ascontiguousarray(idx, np.int32)
    b, c, npoints, nsample = grad_out.shape
    gp = np.zeros((b, c, n), np.float32)
    _lib.oracle_group_points_grad(b, c, n, npoints, nsample, _p(grad_out, _f32p), _p(idx, _i32p), _p(gp, _f32p))
    return gp

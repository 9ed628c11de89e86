"""scn-compatible module surface: same names, constructor arguments, parameter shapes and
error behaviour as the `sparseconvnet` modules composed by models/SparseConvNet.py:59-211."""
import sys

import torch
from torch.nn import Module, Parameter

from . import ops
from .metadata import Metadata


def _counters():
    return sys.modules[__package__]


class SparseConvNetTensor(object):
    """features (N,C) fp32, metadata (shared), spatial_size LongTensor[3]  (Function_test.py:8-11)."""

    def __init__(self, features=None, metadata=None, spatial_size=None):
        self.features = features
        self.metadata = metadata
        self.spatial_size = spatial_size

    def get_spatial_locations(self, spatial_size=None):
        """(N,4) LongTensor [x,y,z,batch] of the active sites, in row order (CPU, like upstream)."""
        size = int((self.spatial_size if spatial_size is None else spatial_size)[0])
        k = self.metadata.levels[size].ukeys
        out = torch.stack([(k >> 32) & 0xFFFF, (k >> 16) & 0xFFFF, k & 0xFFFF, (k >> 48) & 0xFFFF], 1)
        return out.cpu()

    def batch_size(self):
        k = self.metadata.levels[int(self.spatial_size[0])].ukeys
        return int(((k >> 48) & 0xFFFF).max().item()) + 1 if k.numel() else 0

    def to(self, device):
        self.features = self.features.to(device)
        return self

    def type(self, t=None):
        if t:
            self.features = self.features.type(t)
            return self
        return self.features.type()

    def cuda(self):
        self.features = self.features.cuda()
        return self

    def __repr__(self):
        return "SparseConvNetTensor<<features=%s,spatial_size=%s>>" % (
            tuple(self.features.shape) if self.features is not None else None,
            self.spatial_size.tolist() if self.spatial_size is not None else None)


def _size(t):
    return int(t.spatial_size[0])


def _cube(spatial_size, dimension):
    if isinstance(spatial_size, torch.Tensor):
        v = [int(s) for s in spatial_size.tolist()]
    elif isinstance(spatial_size, (list, tuple)):
        v = [int(s) for s in spatial_size]
    else:
        v = [int(spatial_size)] * dimension
    if len(v) != dimension or any(s != v[0] for s in v):
        raise NotImplementedError("b200scn supports cubic spatial sizes only, got %s" % (v,))
    return torch.LongTensor(v)


def _check3(dimension):
    if dimension != 3:
        raise NotImplementedError("b200scn implements dimension 3 only (the reference uses 3 everywhere)")


# Residual blocks (scn.UNet / FullyConvolutionalNet with residual_blocks=True) are ConcatTable(skip, Sequential(..., conv))
# followed by AddTable.  With fusion on (default) Sequential.forward recognises that pair and lets the last convolution
# add the skip branch in its epilogue (one read of the addend instead of a separate read-read-write pass); the module
# tree, parameter names and results are unchanged.
_fuse_residual = [True]


def set_fusion(on):
    _fuse_residual[0] = bool(on)


def _fusable_residual(m, nxt):
    if not (_fuse_residual[0] and type(m) is ConcatTable and type(nxt) is AddTable and len(m._modules) == 2):
        return False
    branch = m._modules["1"]
    if type(branch) is not Sequential or len(branch._modules) == 0:
        return False
    last = list(branch._modules.values())[-1]
    return type(last) is SubmanifoldConvolution and last.bias is None


def _bn_feeds_conv(m, nxt):
    """BatchNorm(Leaky)ReLU directly followed by a convolution: its output is consumed by tensor-core operand fetches only."""
    return isinstance(m, BatchNormalization) and isinstance(nxt, (SubmanifoldConvolution, Convolution, Deconvolution,
                                                                   NetworkInNetwork))


def _fusable_join(m, nxt):
    """ConcatTable(Identity, Sequential(BatchNorm, ...)) + JoinTable -- the down/up path of scn.UNet
    (models/SparseConvNet.py:126-140): x feeds the BatchNorm AND the joined output."""
    if not (_fuse_residual[0] and type(m) is ConcatTable and type(nxt) is JoinTable and len(m._modules) == 2):
        return False
    branch = m._modules["1"]
    if type(m._modules["0"]) is not Identity or type(branch) is not Sequential or len(branch._modules) < 2:
        return False
    return isinstance(list(branch._modules.values())[0], BatchNormalization)


def _run_modules(mods, input):
    """Sequential.forward over a list of modules, with the fusion peepholes (module tree and results unchanged)."""
    i = 0
    while i < len(mods):
        m = mods[i]
        if i + 1 < len(mods) and _fusable_residual(m, mods[i + 1]):
            branch = list(m._modules["1"]._modules.values())
            y = input
            if isinstance(branch[0], BatchNormalization) and len(branch) > 1 and input.features.requires_grad:
                # x feeds the BatchNorm AND the skip: one function returns both, so that backward sums the two
                # gradients inside the BatchNorm backward kernel
                y = branch[0](input, feeds_conv=_bn_feeds_conv(branch[0], branch[1]), want_alias=True)   # (hooks fire)
                input = y.skip_alias
                rest = list(enumerate(branch[:-1]))[1:]
            else:
                rest = list(enumerate(branch[:-1]))
            skip = m._modules["0"](input)
            for j, mod in rest:
                y = mod(y, feeds_conv=True) if _bn_feeds_conv(mod, branch[j + 1]) else mod(y)
            input = branch[-1](y, addend=skip.features)
            i += 2
            continue
        if i + 1 < len(mods) and _fusable_join(m, mods[i + 1]) and input.features.requires_grad:
            # same for the U-Net's down/up path: the gradient that comes back through the joined copy of x is summed inside
            # the BatchNorm backward kernel instead of by an autograd add over the whole tensor (6 per step in the m=32 UNet)
            branch = list(m._modules["1"]._modules.values())
            y = branch[0](input, feeds_conv=_bn_feeds_conv(branch[0], branch[1]), want_alias=True)
            deep = _run_modules(branch[1:], y)
            input = mods[i + 1]([m._modules["0"](y.skip_alias), deep])
            i += 2
            continue
        if i + 1 < len(mods) and _bn_feeds_conv(m, mods[i + 1]):
            input = m(input, feeds_conv=True)
        else:
            input = m(input)
        i += 1
    return input


class Sequential(torch.nn.Sequential):
    def add(self, module):
        self._modules[str(len(self._modules))] = module
        return self

    def forward(self, input):
        return _run_modules(list(self._modules.values()), input)

    def input_spatial_size(self, out_size):
        for m in reversed(self._modules.values()):
            out_size = m.input_spatial_size(out_size)
        return out_size


class Identity(Module):
    def forward(self, input):
        return input

    def input_spatial_size(self, out_size):
        return out_size


class ConcatTable(Sequential):
    def forward(self, input):
        return [module(input) for module in self._modules.values()]

    def input_spatial_size(self, out_size):
        return self._modules["0"].input_spatial_size(out_size)


class AddTable(Module):
    def forward(self, input):
        out = SparseConvNetTensor(None, input[0].metadata, input[0].spatial_size)
        tok = ops._p0("add", "torch add (AddTable)", 12.0 * input[0].features.nelement() * (len(input) - 1), 0, 0.0, 0.0)
        out.features = sum(i.features for i in input)
        ops._p1(tok)
        return out

    def input_spatial_size(self, out_size):
        return out_size


class JoinTable(Module):
    def forward(self, input):
        out = SparseConvNetTensor(None, input[0].metadata, input[0].spatial_size)
        tok = ops._p0("join", "torch cat (JoinTable)", 8.0 * sum(i.features.nelement() for i in input), 0, 0.0, 0.0)
        out.features = torch.cat([i.features for i in input], 1)
        ops._p1(tok)
        return out

    def input_spatial_size(self, out_size):
        return out_size


class InputLayer(Module):
    """scn.InputLayer(dimension, spatial_size, mode=3); forward([coords, feats(, batch_size)]).
    coords: (N,3|4) integer tensor on ANY device (the reference keeps it on the CPU, train.py:58),
    last column = sample index; feats (N,C) float32 on the GPU.  mode 1 last / 2 first / 3 sum / 4 mean."""

    def __init__(self, dimension, spatial_size, mode=3):
        Module.__init__(self)
        _check3(dimension)
        self.dimension = dimension
        self.spatial_size = _cube(spatial_size, dimension)
        self.mode = mode
        self.device = None

    def forward(self, input):
        coords, feats = input[0], input[1]
        ops.recover_deferred()   # (a backward pass that died half-way in deferred-dW mode leaves the streams unjoined)
        if self.mode not in (1, 2, 3, 4):
            raise NotImplementedError("InputLayer mode %r (supported: 1 last, 2 first, 3 sum, 4 mean)" % (self.mode,))
        if not feats.is_cuda:
            raise RuntimeError("b200scn InputLayer: features must live on a CUDA device (no CPU fallback)")
        assert coords.size(0) == feats.size(0), "coords and features disagree on the number of points"
        md = Metadata(self.dimension)
        level = md.build_input(coords, int(self.spatial_size[0]), self.mode, feats.device)
        out = SparseConvNetTensor(None, md, self.spatial_size)
        out.features = ops.InputFeaturesFn.apply(feats, md, level.n)
        return out

    def input_spatial_size(self, out_size):
        return out_size


class OutputLayer(Module):
    def __init__(self, dimension):
        Module.__init__(self)
        self.dimension = dimension

    def forward(self, input):
        return ops.OutputFeaturesFn.apply(input.features, input.metadata)

    def input_spatial_size(self, out_size):
        return out_size


class SceneMeanPooling(Module):
    """OutputLayer followed by the per-scene mean over points that the reference's heads take
    (SparseConvBase_.postProcessing, models/SparseConvNet.py:20-26; models/MultiLabelContrastive.py:35-40), as one fused
    op on the level-0 voxel features: forward(x[, batch_size]) -> (B, C).  Not part of upstream scn: an optional
    replacement for `OutputLayer` + the Python loop of torch.mean when only the pooled features are needed."""

    def forward(self, input, batch_size=None):
        md = input.metadata
        level = md.levels[_size(input)]
        if md.count is None or md.count.shape[0] != input.features.shape[0]:
            raise ValueError("SceneMeanPooling needs the level-0 (InputLayer resolution) tensor")
        B = int(batch_size) if batch_size is not None else input.batch_size()
        return ops.SceneMeanFn.apply(input.features, md, level, B)


class _ConvBase(Module):
    def _init_weight(self, dimension, nIn, nOut, filter_size, bias, groups):
        _check3(dimension)
        if groups != 1:
            raise NotImplementedError("groups != 1 (the reference never passes groups)")
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size = int(filter_size)
        self.filter_volume = self.filter_size ** dimension
        std = (2.0 / nIn / self.filter_volume) ** 0.5
        self.weight = Parameter(torch.Tensor(self.filter_volume, 1, nIn, nOut).normal_(0, std))
        if bias:
            self.bias = Parameter(torch.Tensor(nOut).zero_())
        else:
            self.bias = None

    def _w(self):
        return self.weight.view(self.filter_volume, self.nIn, self.nOut)

    def _finish(self, input, features, spatial_size, nrules):
        if self.bias is not None:
            features = features + self.bias
        c = _counters()
        c._add_madds(nrules, self.nIn * self.nOut)
        c.forward_pass_hidden_states += features.nelement()
        out = SparseConvNetTensor(features, input.metadata, spatial_size)
        return out


class SubmanifoldConvolution(_ConvBase):
    """scn.SubmanifoldConvolution(dimension, nIn, nOut, filter_size, bias, groups=1); weight (27,1,nIn,nOut)."""

    def __init__(self, dimension, nIn, nOut, filter_size, bias, groups=1):
        Module.__init__(self)
        if int(filter_size) != 3:
            raise NotImplementedError("SubmanifoldConvolution: filter_size 3 only (the reference uses 3)")
        self._init_weight(dimension, nIn, nOut, filter_size, bias, groups)

    def forward(self, input, addend=None):
        """addend: optional (N, nOut) features added to the result in the kernel's epilogue (fused residual add)."""
        assert input.features.nelement() == 0 or input.features.size(1) == self.nIn, (self.nIn, self.nOut, input)
        level = input.metadata.levels[_size(input)]
        feats = ops.SubmanifoldConvFn.apply(input.features, self._w(), level, addend, getattr(input, "tf32_rounded", False))
        return self._finish(input, feats, input.spatial_size, level)

    def input_spatial_size(self, out_size):
        return out_size

    def __repr__(self):
        return "SubmanifoldConvolution %d->%d C3" % (self.nIn, self.nOut)


class Convolution(_ConvBase):
    """scn.Convolution(dimension, nIn, nOut, filter_size, filter_stride, bias); size == stride (2 or 4)."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        Module.__init__(self)
        if int(filter_size) != int(filter_stride):
            raise NotImplementedError("Convolution: filter_size == filter_stride only (as in the reference)")
        self.filter_stride = int(filter_stride)
        self._init_weight(dimension, nIn, nOut, filter_size, bias, groups)

    def forward(self, input):
        assert input.features.nelement() == 0 or input.features.size(1) == self.nIn
        s = self.filter_size
        size = _size(input)
        csize = (size - s) // s + 1
        assert (csize - 1) * s + s == size, "spatial size %d not compatible with size/stride %d" % (size, s)
        down = input.metadata.get_down(size, s)
        feats = ops.ConvolutionFn.apply(input.features, self._w(), down)
        return self._finish(input, feats, torch.LongTensor([csize] * 3), down.fine.n)

    def input_spatial_size(self, out_size):
        return (out_size - 1) * self.filter_stride + self.filter_size

    def __repr__(self):
        return "Convolution %d->%d C%d/%d" % (self.nIn, self.nOut, self.filter_size, self.filter_stride)


class Deconvolution(_ConvBase):
    """scn.Deconvolution(dimension, nIn, nOut, filter_size, filter_stride, bias) (decoder of scn.UNet)."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        Module.__init__(self)
        if int(filter_size) != int(filter_stride):
            raise NotImplementedError("Deconvolution: filter_size == filter_stride only")
        self.filter_stride = int(filter_stride)
        self._init_weight(dimension, nIn, nOut, filter_size, bias, groups)

    def forward(self, input):
        assert input.features.nelement() == 0 or input.features.size(1) == self.nIn
        s = self.filter_size
        down = input.metadata.get_up(_size(input), s)
        feats = ops.DeconvolutionFn.apply(input.features, self._w(), down)
        return self._finish(input, feats, torch.LongTensor([down.fine.size] * 3), down.fine.n)

    def input_spatial_size(self, out_size):
        return (out_size - self.filter_size) // self.filter_stride + 1

    def __repr__(self):
        return "Deconvolution %d->%d C%d/%d" % (self.nIn, self.nOut, self.filter_size, self.filter_stride)


class UnPooling(Module):
    def __init__(self, dimension, pool_size, pool_stride):
        Module.__init__(self)
        _check3(dimension)
        if int(pool_size) != int(pool_stride):
            raise NotImplementedError("UnPooling: pool_size == pool_stride only")
        self.pool_size, self.pool_stride = int(pool_size), int(pool_stride)

    def forward(self, input):
        down = input.metadata.get_up(_size(input), self.pool_size)
        out = SparseConvNetTensor(None, input.metadata, torch.LongTensor([down.fine.size] * 3))
        out.features = ops.UnPoolingFn.apply(input.features, down)
        return out

    def input_spatial_size(self, out_size):
        return (out_size - self.pool_size) // self.pool_stride + 1


class MaxPooling(Module):
    """scn.MaxPooling(dimension, pool_size, pool_stride, nFeaturesToDrop=0), pool_size == pool_stride
    (models/projector/components.py:78-100)."""

    def __init__(self, dimension, pool_size, pool_stride, nFeaturesToDrop=0):
        Module.__init__(self)
        _check3(dimension)
        if int(pool_size) != int(pool_stride):
            raise NotImplementedError("MaxPooling: pool_size == pool_stride only")
        if nFeaturesToDrop:
            raise NotImplementedError("MaxPooling: nFeaturesToDrop != 0")
        self.pool_size, self.pool_stride = int(pool_size), int(pool_stride)

    def forward(self, input):
        s = self.pool_size
        size = _size(input)
        csize = (size - s) // s + 1
        assert (csize - 1) * s + s == size, "spatial size %d not compatible with size/stride %d" % (size, s)
        down = input.metadata.get_down(size, s)
        out = SparseConvNetTensor(None, input.metadata, torch.LongTensor([csize] * 3))
        out.features = ops.MaxPoolingFn.apply(input.features, down)
        return out

    def input_spatial_size(self, out_size):
        return (out_size - 1) * self.pool_stride + self.pool_size


class SparseToDense(Module):
    """scn.SparseToDense(dimension, nPlanes): (N, C) features -> dense (B, C, S, S, S) (Function_test.py:46)."""

    def __init__(self, dimension, nPlanes):
        Module.__init__(self)
        _check3(dimension)
        self.nPlanes = nPlanes

    def forward(self, input):
        level = input.metadata.levels[_size(input)]
        return ops.SparseToDenseFn.apply(input.features, level, input.batch_size())

    def input_spatial_size(self, out_size):
        return out_size


class NetworkInNetwork(Module):
    def __init__(self, nIn, nOut, bias):
        Module.__init__(self)
        self.nIn, self.nOut = nIn, nOut
        std = (2.0 / nIn) ** 0.5
        self.weight = Parameter(torch.Tensor(nIn, nOut).normal_(0, std))
        if bias:
            self.bias = Parameter(torch.Tensor(nOut).zero_())
        else:
            self.bias = None

    def forward(self, input):
        assert input.features.nelement() == 0 or input.features.size(1) == self.nIn
        feats = ops.NetworkInNetworkFn.apply(input.features, self.weight)
        if self.bias is not None:
            feats = feats + self.bias
        c = _counters()
        c._add_madds(input.features.size(0), self.nIn * self.nOut)
        c.forward_pass_hidden_states += feats.nelement()
        return SparseConvNetTensor(feats, input.metadata, input.spatial_size)

    def input_spatial_size(self, out_size):
        return out_size

    def __repr__(self):
        return "NetworkInNetwork %d->%d" % (self.nIn, self.nOut)


class BatchNormalization(Module):
    """scn.BatchNormalization(nPlanes, eps=1e-4, momentum=0.9, affine=True, leakiness=1)."""

    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, affine=True, leakiness=1):
        Module.__init__(self)
        self.nPlanes, self.eps, self.momentum, self.affine, self.leakiness = nPlanes, eps, momentum, affine, leakiness
        self.register_buffer("running_mean", torch.Tensor(nPlanes).fill_(0))
        self.register_buffer("running_var", torch.Tensor(nPlanes).fill_(1))
        if affine:
            self.weight = Parameter(torch.Tensor(nPlanes).fill_(1))
            self.bias = Parameter(torch.Tensor(nPlanes).fill_(0))
        else:
            self.register_buffer("weight", torch.Tensor(nPlanes).fill_(1))
            self.register_buffer("bias", torch.Tensor(nPlanes).fill_(0))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # accept Lua-era camelCase buffer names as well (checkpoint compatibility)
        for old, new in (("runningMean", "running_mean"), ("runningVar", "running_var")):
            if prefix + old in state_dict and prefix + new not in state_dict:
                state_dict[prefix + new] = state_dict.pop(prefix + old)
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def forward(self, input, feeds_conv=False, want_alias=False):
        """feeds_conv: set by Sequential when the next module is a convolution -- on the TF32 path the output is then
        written rounded to the nearest TF32 (the tensor core would truncate it when it fetches the operand).
        want_alias: the result also carries `.skip_alias`, a tensor with the INPUT's features for the residual skip;
        gradients flowing back into it are summed with the BatchNorm's own input gradient in one kernel
        (ops.BatchNormSkipFn) instead of by a separate autograd add."""
        assert input.features.nelement() == 0 or input.features.size(1) == self.nPlanes
        out = SparseConvNetTensor(None, input.metadata, input.spatial_size)
        args = (input.features, self.weight, self.bias, self.running_mean, self.running_var, self.eps, self.momentum,
                self.training, self.leakiness, feeds_conv)
        out.tf32_rounded = bool(feeds_conv) and ops.get_precision() == "tf32"   # consumers need not round again
        if want_alias:
            out.features, xa = ops.BatchNormSkipFn.apply(*args)
            out.skip_alias = SparseConvNetTensor(xa, input.metadata, input.spatial_size)
        else:
            out.features = ops.BatchNormFn.apply(*args)
        return out

    def input_spatial_size(self, out_size):
        return out_size

    def __repr__(self):
        return "BatchNorm(%d,eps=%g,momentum=%g,affine=%s,leakiness=%g)" % (
            self.nPlanes, self.eps, self.momentum, self.affine, self.leakiness)


class BatchNormReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9):
        BatchNormalization.__init__(self, nPlanes, eps, momentum, True, 0)

    def __repr__(self):
        return "BatchNormReLU(%d,eps=%g,momentum=%g,affine=True)" % (self.nPlanes, self.eps, self.momentum)


class BatchNormLeakyReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, leakiness=0.333):
        BatchNormalization.__init__(self, nPlanes, eps, momentum, True, leakiness)

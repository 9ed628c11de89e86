"""Run-to-run spread of every TF32 op on fixed inputs (diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "3d-weakly-supervised-semantic-segmentation_b200"))
import torch
import sparseconvnet as scn
from sparseconvnet import ops
from b200scn_synth import make_batch

def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))

def spread(name, fn, reps=5):
    ref = fn()
    torch.cuda.synchronize()
    s = 0.0
    for _ in range(reps):
        out = fn()
        torch.cuda.synchronize()
        s = max(s, rel(out, ref))
    print("%-40s spread %.1e   bitwise %s" % (name, s, s == 0.0))

scn.set_precision("tf32")
coords, feats, _ = make_batch([0, 1], 40, n_points=60000)
x = scn.InputLayer(3, 4096, mode=4)([coords, feats.cuda()])
md = x.metadata
l0 = md.levels[4096]
torch.manual_seed(0)
f16 = torch.randn(l0.n, 16, device="cuda"); f32 = torch.randn(l0.n, 32, device="cuda")
w = torch.randn(27, 16, 16, device="cuda") * 0.1
spread("tiled subm 16->16 L0 (n=%d)" % l0.n, lambda: ops.subm_conv(f16, l0, ops.GemmWeight(w)))
spread("gather subm 16->16 L0", lambda: ops.gather_conv(f16, l0.subm_map(), l0.n, 27, ops.GemmWeight(w), rules=l0))
conv = scn.Convolution(3, 16, 32, 2, 2, False).cuda()
t16 = scn.SparseConvNetTensor(f16, md, x.spatial_size)
y = conv(t16)
l1n = y.features.shape[0]
spread("Convolution 16->32 s2 (n1=%d)" % l1n, lambda: conv(t16).features)
deconv = scn.Deconvolution(3, 32, 16, 2, 2, False).cuda()
t32 = scn.SparseConvNetTensor(torch.randn(l1n, 32, device="cuda"), md, y.spatial_size)
spread("Deconvolution 32->16 s2", lambda: deconv(t32).features)
nin = scn.NetworkInNetwork(32, 16, False).cuda()
t32b = scn.SparseConvNetTensor(f32, md, x.spatial_size)
spread("NetworkInNetwork 32->16", lambda: nin(t32b).features)
bn = scn.BatchNormReLU(16).cuda()
spread("BatchNormReLU 16 (feeds_conv rounding)", lambda: bn(t16, feeds_conv=True).features)
spread("BatchNormReLU 16", lambda: bn(t16).features)
stem = scn.SubmanifoldConvolution(3, 3, 16, 3, False).cuda()
spread("stem SubM 3->16", lambda: stem(x).features)
# backward pieces
g16 = torch.randn(l0.n, 16, device="cuda")
pin, pout, offs = l0.subm_pairs_ordered(l0.tile_plan(ops._halo["hcap"]).perm)
spread("pair_dw 16x16 L0", lambda: ops.pair_dw(f16, g16, pin, pout, offs, 27, l0.n))
def bn_bwd():
    xin = f16.detach().requires_grad_(True)
    out = bn(scn.SparseConvNetTensor(xin, md, x.spatial_size)).features
    out.backward(g16)
    return xin.grad
spread("BatchNormReLU backward dx", bn_bwd)
def conv_bwd():
    xin = f16.detach().requires_grad_(True)
    conv.weight.grad = None
    out = conv(scn.SparseConvNetTensor(xin, md, x.spatial_size)).features
    out.backward(torch.ones_like(out))
    return torch.cat([xin.grad.reshape(-1), conv.weight.grad.reshape(-1)])
spread("Convolution backward (dx, dw)", conv_bwd)
def deconv_bwd():
    xin = t32.features.detach().requires_grad_(True)
    deconv.weight.grad = None
    out = deconv(scn.SparseConvNetTensor(xin, md, y.spatial_size)).features
    out.backward(torch.ones_like(out))
    return torch.cat([xin.grad.reshape(-1), deconv.weight.grad.reshape(-1)])
spread("Deconvolution backward (dx, dw)", deconv_bwd)

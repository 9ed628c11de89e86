import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/3d-weakly-supervised-semantic-segmentation_b200'); sys.path.insert(0,'/root/repo/tests')
import torch
import sparseconvnet as scn
from oracle import scn_oracle as ref
from b200scn_synth import build_encoder, make_batch
from _util import rel_err, copy_params
torch.manual_seed(0)
coords, feats, _ = make_batch([0,1], 20, n_points=20000)
nr = build_encoder(ref,"SparseConvUNet",16,1,False)
ng = build_encoder(scn,"SparseConvUNet",16,1,False)
copy_params(nr,ng); ng.cuda()
outs_g, outs_r = [], []
def hk(store):
    def f(mod, inp, out):
        store.append((inp[0].features.detach().cpu(), out.features.detach().cpu()))
    return f
for m in ng.modules():
    if isinstance(m, scn.BatchNormalization): m.register_forward_hook(hk(outs_g))
for m in nr.modules():
    if isinstance(m, ref.BatchNormalization): m.register_forward_hook(hk(outs_r))
def run_g():
    fg=feats.clone().cuda().requires_grad_(True)
    og=ng([coords,fg])
    torch.manual_seed(1); go=torch.randn(og.shape)/og.shape[0]
    for p in ng.parameters(): p.grad=None
    og.backward(go.cuda())
    return og.detach().cpu(), fg.grad.cpu(), [p.grad.cpu().clone() for p in ng.parameters()]
o1,g1,w1 = run_g()
n_bn = len(outs_g)
o2,g2,w2 = run_g()
print('gpu run-to-run: logits %.2e input grad %.2e w0 grad %.2e' % (rel_err(o1,o2), rel_err(g1,g2), rel_err(w1[0],w2[0])))
fr=feats.clone().requires_grad_(True)
o_r=nr([coords,fr]); torch.manual_seed(1); go=torch.randn(o_r.shape)/o_r.shape[0]; o_r.backward(go)
print('gpu vs ref: logits %.2e input grad %.2e' % (rel_err(o1,o_r), rel_err(g1,fr.grad)))
tot=0
for i in range(n_bn):
    (xi_g, yg), (xi_r, yr) = outs_g[i], outs_r[i]
    mism = ((yg>0)!=(yr>0))
    tot += int(mism.sum())
    print(i, tuple(yg.shape), 'in err %.2e out err %.2e mask mismatches %d  max|y| at mismatch %.2e' % (rel_err(xi_g,xi_r), rel_err(yg,yr), int(mism.sum()), float(torch.maximum(yg.abs(),yr.abs())[mism].max()) if mism.any() else 0))
print('total flips', tot)

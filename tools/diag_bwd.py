"""Diagnostic: per-module gradient comparison GPU vs oracle (and GPU run-to-run)."""
import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/3d-weakly-supervised-semantic-segmentation_b200'); sys.path.insert(0,'/root/repo/tests')
import torch
import sparseconvnet as scn
from oracle import scn_oracle as ref
from b200scn_synth import build_encoder, make_batch
from _util import rel_err, copy_params
smooth = len(sys.argv) > 1 and sys.argv[1] == 'smooth'
torch.manual_seed(0)
coords, feats, _ = make_batch([0,1], 20, n_points=20000)
nr = build_encoder(ref,"SparseConvUNet",16,1,False)
ng = build_encoder(scn,"SparseConvUNet",16,1,False)
if smooth:
    for net in (nr, ng):
        for m in net.modules():
            if hasattr(m,'leakiness'): m.leakiness = 1.0
copy_params(nr,ng); ng.cuda()
def instrument(net, ns, store):
    names = {m: n for n, m in net.named_modules()}
    def hk(mod, inp, out):
        if hasattr(out, 'features') and out.features.requires_grad:
            nm = names[mod] + ':' + type(mod).__name__
            out.features.register_hook(lambda g, nm=nm: store.append((nm, g.detach().cpu().clone())))
    for m in net.modules():
        if len(list(m.children())) == 0:
            m.register_forward_hook(hk)
sg, sr = [], []
instrument(ng, scn, sg); instrument(nr, ref, sr)
def run(net, dev, store):
    store.clear()
    f=feats.clone().to(dev).requires_grad_(True)
    o=net([coords,f]); torch.manual_seed(1); go=torch.randn(o.shape)/o.shape[0]
    for p in net.parameters(): p.grad=None
    o.backward(go.to(dev))
    return list(store), f.grad.cpu()
g1, fg1 = run(ng,'cuda',sg)
g2, fg2 = run(ng,'cuda',sg)
r1, fr1 = run(nr,'cpu',sr)
print('input grad gpu-vs-ref %.2e  gpu run-to-run %.2e' % (rel_err(fg1,fr1), rel_err(fg1,fg2)))
for (n1,a),(n2,b),(n3,c) in zip(g1,g2,r1):
    assert n1==n3, (n1,n3)
    print('%-40s %-18s gpu-vs-ref %.2e  run-to-run %.2e' % (n1, tuple(a.shape), rel_err(a,c), rel_err(a,b)))

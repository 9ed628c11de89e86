"""Run-to-run spread of a small two-level U-Net, per precision and per stage (diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "3d-weakly-supervised-semantic-segmentation_b200"))
import torch
import sparseconvnet as scn
from sparseconvnet import ops
from b200scn_synth import make_batch

def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))

coords, feats, _ = make_batch([0, 1], 40, n_points=60000)
feats = feats.cuda()
for prec, cache in (("tf32", True), ("tf32", False)):
    scn.set_precision(prec)
    scn.set_weight_cache(cache)
    print("weight cache", cache)
    torch.manual_seed(0)
    net = scn.Sequential(scn.InputLayer(3, 4096, mode=4), scn.SubmanifoldConvolution(3, 3, 16, 3, False),
                         scn.UNet(3, 2, [16, 32], True), scn.BatchNormReLU(16), scn.OutputLayer(3)).cuda()
    for mod in net.modules():
        if hasattr(mod, "leakiness"):
            mod.leakiness = 1.0
    def run():
        for p in net.parameters():
            p.grad = None
        f = feats.detach().requires_grad_(True)
        out = net([coords, f])
        (out * out).mean().backward()
        torch.cuda.synchronize()
        return out.detach().clone(), f.grad.clone(), {n: p.grad.clone() for n, p in net.named_parameters()}
    o0, g0, w0 = run()
    prev = o0
    for rep in range(3):
        o1, g1, w1 = run()
        print("   forward vs previous run %.1e" % rel(o1, prev))
        prev = o1
        worst = max(w0, key=lambda n: rel(w1[n], w0[n]))
        print(prec, "rep", rep, "forward %.1e  input grad %.1e  worst weight grad %.1e (%s)" % (rel(o1, o0), rel(g1, g0), rel(w1[worst], w0[worst]), worst))
    names = list(w0)
    print("   per-parameter spread:", ", ".join("%s %.0e" % (n.replace("weight", "w").replace("bias", "b"), rel(w1[n], w0[n])) for n in names[:24]))

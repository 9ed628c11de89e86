"""Tiled (conv_halo.cu) vs gather (conv_tc.cu) submanifold kernels over a channel sweep: max relative difference."""
import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/3d-weakly-supervised-semantic-segmentation_b200'); sys.path.insert(0, '/root/repo/tests')
import torch
import sparseconvnet as scn
from sparseconvnet import ops
from _util import random_cloud
scn.set_precision("tf32")
coords, feats = random_cloud(7, 3000, 24, 2)
x = scn.InputLayer(3, 4096, mode=4)([coords, feats.cuda()])
level = x.metadata.levels[4096]
bad = []
for cin in [int(v) for v in os.environ.get("SWEEP_CIN", "32,96,128,160,192,224,256,320,384").split(",")]:
    for cout in [int(v) for v in os.environ.get("SWEEP_COUT", "32,64,96,128,192,384").split(",")]:
        torch.manual_seed(cin * 1000 + cout)
        f = torch.randn(level.n, cin, device='cuda')
        w = torch.randn(27, cin, cout, device='cuda') * 0.1
        gw = ops.GemmWeight(w)
        scn.set_tiled("off")
        ref = ops.subm_conv(f, level, gw)
        scn.set_tiled("on")
        errs = []
        for rep in range(3):
            out = ops.subm_conv(f, level, gw)
            errs.append(float((out - ref).norm() / ref.norm()))
        flag = "" if max(errs) < 1e-5 else "  <-- MISMATCH"
        if flag: bad.append((cin, cout))
        print("cin %3d cout %3d  rel diff %s%s" % (cin, cout, " ".join("%.1e" % e for e in errs), flag))
print("mismatches:", bad)

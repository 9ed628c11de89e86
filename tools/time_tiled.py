"""Time the tiled submanifold kernel (conv_halo.cu) per benchmark level, with and without operand rounding, and check it
against the gather kernel (relative error).   python tools/time_tiled.py"""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/3d-weakly-supervised-semantic-segmentation_b200')
import torch
import sparseconvnet as scn
from sparseconvnet import ops
from b200scn_synth import make_batch
scn.set_precision("tf32")
coords, feats, _ = make_batch(list(range(5)), 50)
x = scn.InputLayer(3, 4096, mode=4)([coords, feats.cuda()])
md = x.metadata
shapes = [(32, 32, 0), (64, 32, 0), (64, 64, 1), (128, 64, 1), (96, 96, 2), (192, 96, 2), (128, 128, 3), (256, 128, 3)]
tot = 0.0
for cin, cout, lvl in shapes:
    level = md.levels[4096 >> lvl]
    f = torch.randn(level.n, cin, device='cuda')
    w = torch.randn(27, cin, cout, device='cuda') * 0.1
    gw = ops.GemmWeight(w)
    ref = ops.gather_conv(f, level.subm_map(), level.n, 27, gw, rules=level)
    line = "level %d n %7d %3d->%3d" % (lvl, level.n, cin, cout)
    for ra in (True, False):
        for _ in range(3):
            y = ops.subm_conv(f, level, gw, round_a=ra)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            y = ops.subm_conv(f, level, gw, round_a=ra)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 50
        err = ((y - ref).norm() / ref.norm()).item()
        line += "   round_a=%d %7.1f us rel %.1e" % (ra, us, err)
    print(line)

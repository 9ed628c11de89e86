"""Knockout timing of the pair-list weight-gradient kernel (b200scn_set_option "dw_dbg": 1 no MMA issue, 2 no gathers).
Results are wrong by construction; only the time matters.   python tools/dw_knockout.py"""
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
for ca, cg, lvl in [(32, 32, 0), (64, 64, 1), (128, 128, 3)]:
    level = md.levels[4096 >> lvl]
    a = torch.randn(level.n, ca, device='cuda'); g = torch.randn(level.n, cg, device='cuda')
    pin, pout, offs = level.subm_pairs_ordered(level.tile_plan(ops._halo["hcap"]).perm)
    line = "level %d %3dx%3d pairs %8d:" % (lvl, ca, cg, int(offs[-1]))
    for name, dbg in (("full", 0), ("no MMA", 1), ("no gathers", 2), ("neither", 3)):
        scn.set_option("dw_dbg", dbg)
        for _ in range(3):
            ops.pair_dw(a, g, pin, pout, offs, 27, level.n)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.pair_dw(a, g, pin, pout, offs, 27, level.n)
        e1.record()
        torch.cuda.synchronize()
        line += "  %s %6.1f us" % (name, e0.elapsed_time(e1) * 100)
    scn.set_option("dw_dbg", 0)
    print(line)

"""Time the pair-list weight-gradient kernel per benchmark level: per-offset chunks vs row-block-aligned segments
(b200scn_pair_dw vs b200scn_pair_dw_blocked), and check they agree.   python tools/time_dw.py"""
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


def t(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


for ca, cg, lvl in [(32, 32, 0), (64, 32, 0), (64, 64, 1), (128, 64, 1), (96, 96, 2), (192, 96, 2), (128, 128, 3), (256, 128, 3)]:
    level = md.levels[4096 >> lvl]
    torch.manual_seed(lvl)
    a = torch.randn(level.n, ca, device='cuda'); g = torch.randn(level.n, cg, device='cuda')
    perm = level.tile_plan(ops._halo["hcap"]).perm
    pin, pout, offs, (blk, nblk) = level.subm_pairs_blocked(perm)
    ref = torch.einsum("pa,pg->ag", a[pin[:int(offs[1])].long()].double(), g[pout[:int(offs[1])].long()].double())   # offset 0, fp64
    line = "level %d n %7d %3dx%3d  pairs %8d" % (lvl, level.n, ca, cg, int(offs[-1]))
    for pairs in (32, 64):
        scn.set_option("dw_pairs", pairs)
        d0 = ops.pair_dw(a, g, pin, pout, offs, 27, level.n)
        d1 = ops.pair_dw_blocked(a, g, pin, pout, blk, nblk, 27)
        err = ((d0 - d1).norm() / d0.norm()).item()
        err0 = ((d0[0].double() - ref).norm() / ref.norm()).item()
        t0 = t(lambda: ops.pair_dw(a, g, pin, pout, offs, 27, level.n))
        t1 = t(lambda: ops.pair_dw_blocked(a, g, pin, pout, blk, nblk, 27))
        line += "  | stage %d: chunks %7.1f us  row blocks %7.1f us  diff %.0e  vs fp64 %.1e" % (pairs, t0, t1, err, err0)
    scn.set_option("dw_pairs", 32)
    d0 = ops.pair_dw(a, g, pin, pout, offs, 27, level.n)
    dt = ops.subm_dw_tiled(a, g, level)
    if dt is not None:
        errt = ((dt - d0).norm() / d0.norm()).item()
        errt0 = ((dt[0].double() - ref).norm() / ref.norm()).item()
        tt = t(lambda: ops.subm_dw_tiled(a, g, level))
        line += "  | tile-stationary %7.1f us  diff %.0e  vs fp64 %.1e" % (tt, errt, errt0)
    print(line)

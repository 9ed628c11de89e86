"""Weight-gradient kernels side by side on the cfg3 grids: pair-list kernel (Morton-ordered lists) vs tile-stationary."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "3d-weakly-supervised-semantic-segmentation_b200"))
import torch
import sparseconvnet as scn
from sparseconvnet import ops
from b200scn_synth import make_batch

scn.set_precision("tf32")
coords, feats, _ = make_batch(list(range(5)), 50)
x = scn.InputLayer(3, 4096, mode=4)([coords, feats.cuda()])
md = x.metadata
def t(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for level, ca, cg in [(0, 32, 32), (0, 64, 32), (1, 64, 64), (1, 128, 64), (2, 96, 96), (2, 192, 96), (3, 128, 128), (3, 256, 128)]:
    lvl = md.levels[4096 >> level]
    a = torch.randn(lvl.n, ca, device="cuda"); g = torch.randn(lvl.n, cg, device="cuda")
    plan = lvl.tile_plan(ops._halo["hcap"])
    pin, pout, offs = lvl.subm_pairs_ordered(plan.perm)
    R = sum(lvl.rule_counts())
    tp = t(lambda: ops.pair_dw(a, g, pin, pout, offs, 27, lvl.n))
    ok = ops.subm_dw_tiled(a, g, lvl) is not None
    tt = t(lambda: ops.subm_dw_tiled(a, g, lvl)) if ok else float("nan")
    fl = 2.0 * R * ca * cg
    by = 4.0 * lvl.n * (ca + cg) + 8.0 * R + 4 * 27 * ca * cg
    print("L%d %3dx%-3d n=%7d R=%8d  pair %7.1f us (%5.1f TF/s %5.0f GB/s)   tiled %7.1f us (%5.1f TF/s %5.0f GB/s)  x%.2f" % (
        level, ca, cg, lvl.n, R, tp, fl / tp / 1e6, by / tp / 1e3, tt, fl / tt / 1e6, by / tt / 1e3, tp / tt))

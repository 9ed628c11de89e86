"""Launch a fixed set of hot kernels once each (after a warm-up) for an `ncu --set full` capture:
tiled SubM conv L0 32->32 and L1 64->64, tile-stationary dW L0 32x32 and L1 64x64, pair dW L1 64x64, BN fwd/bwd L0."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "3d-weakly-supervised-semantic-segmentation_b200"))
import torch
import sparseconvnet as scn
from sparseconvnet import ops
from b200scn_synth import make_batch

which = sys.argv[1:] or ["conv", "dw"]
scn.set_precision("tf32")
coords, feats, _ = make_batch(list(range(5)), 50)
x = scn.InputLayer(3, 4096, mode=4)([coords, feats.cuda()])
md = x.metadata
for rep in range(2):          # rep 0 = warm-up (skip with ncu --launch-skip), rep 1 = captured
    for level, c in ((0, 32), (1, 64)):
        lvl = md.levels[4096 >> level]
        torch.manual_seed(level)
        f = torch.randn(lvl.n, c, device="cuda"); g = torch.randn(lvl.n, c, device="cuda")
        w = torch.randn(27, c, c, device="cuda") * 0.1
        if "conv" in which:
            ops.subm_conv(f, lvl, ops.GemmWeight(w))
        if "dw" in which:
            ops.subm_dw_tiled(f, g, lvl)
        if "pair" in which:
            pin, pout, offs = lvl.subm_pairs_ordered(lvl.tile_plan(ops._halo["hcap"]).perm)
            ops.pair_dw(f, g, pin, pout, offs, 27, lvl.n)
    torch.cuda.synchronize()
print("ncu_target done", which, [md.levels[4096 >> i].n for i in range(2)])

"""Knockout timing of the tiled submanifold kernel (conv_halo.cu): the same launch with parts of the pipeline removed
(b200scn_set_option "halo_dbg": 1 no A build, 2 no MMA issue, 4 no halo reads (TMEM stores kept), 8 no TMEM stores (halo reads
kept), 16 no global halo copies, 32 no output stores).  Results are wrong by construction; only the time matters: it shows
which stage the launch is bound by.   python tools/halo_knockout.py [cin cout level]..."""
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
shapes = [(32, 32, 0), (64, 64, 1), (96, 96, 2), (128, 128, 3), (128, 64, 1)]
if len(sys.argv) > 3:
    a = [int(v) for v in sys.argv[1:]]
    shapes = [tuple(a[i:i + 3]) for i in range(0, len(a), 3)]
CASES = [("full", 0), ("no A build", 1), ("no MMA", 2), ("no A build, no MMA", 3), ("no halo LDS", 4), ("no TMEM st", 8),
         ("no global halo copy", 16), ("no output store", 32), ("no build/MMA/copy/store", 1 | 2 | 16 | 32), ("half W bytes", 64), ("half W, no build/MMA/copy/store", 64 | 1 | 2 | 16 | 32),
         ("one CTA/SM", -1), ("one CTA/SM skeleton", -(1 | 2 | 16 | 32))]
for cin, cout, lvl in shapes:
    level = md.levels[4096 >> lvl]
    f = torch.randn(level.n, cin, device='cuda')
    w = torch.randn(27, cin, cout, device='cuda') * 0.1
    gw = ops.GemmWeight(w)
    print("== level %d  n %d  %d -> %d  tiles %d" % (lvl, level.n, cin, cout, (level.n + 127) // 128))
    for name, dbg in CASES:
        scn.set_option("halo_one_cta", 1 if dbg < 0 else 0)
        scn.set_option("halo_dbg", 0 if dbg == -1 else abs(dbg))
        for _ in range(3):
            ops.subm_conv(f, level, gw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.subm_conv(f, level, gw)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        cyc = us * 1e-6 * 1.965e9 * 148 / ((level.n + 127) // 128)
        print("  %-28s %8.1f us   %7.0f SM-cycles/tile" % (name, us, cyc))
    scn.set_option("halo_dbg", 0)
    scn.set_option("halo_one_cta", 0)

"""clock64 timeline of CTA 0 of the tile-stationary weight-gradient kernel (diagnostic; dw_tile.cu DW_TL slots)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "3d-weakly-supervised-semantic-segmentation_b200"))
import torch
import sparseconvnet as scn
from sparseconvnet import ops, _lib
from b200scn_synth import make_batch
ca, cg, lv = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
scn.set_precision("tf32")
coords, feats, _ = make_batch(list(range(5)), 50)
x = scn.InputLayer(3, 4096, mode=4)([coords, feats.cuda()])
lvl = x.metadata.levels[4096 >> lv]
a = torch.randn(lvl.n, ca, device="cuda"); g = torch.randn(lvl.n, cg, device="cuda")
for _ in range(2): ops.subm_dw_tiled(a, g, lvl)
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
fn = _lib.lib.b200scn_debug_dw_timeline; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p]
assert fn(buf.data_ptr()) == 0
ops.subm_dw_tiled(a, g, lvl); torch.cuda.synchronize()
fn(None)
t = buf.cpu().numpy().reshape(64, 16)
t0 = t[0, 0]
print("level", lv, "n", lvl.n, ca, "x", cg)
print("it: load[wait  free  plan  landed  published]   prod0[seen done]   mma[seen done] stages")
for it in range(12):
    r = [int(v - t0) if v else -1 for v in t[it, :9]]
    print("%2d: %7d %7d %7d %7d %7d   | %7d %7d | %7d %7d  st=%d" % (it, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], int(t[it, 9])))

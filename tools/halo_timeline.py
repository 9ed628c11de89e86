"""clock64 timeline of one CTA of the spatially tiled submanifold kernel (diagnostic; conv_halo.cu SCN_TL slots)."""
import sys, ctypes
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/3d-weakly-supervised-semantic-segmentation_b200')
import torch
import sparseconvnet as scn
from sparseconvnet import ops, _lib
from b200scn_synth import make_batch
cin, cout, lvl = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dbg = int(sys.argv[4]) if len(sys.argv) > 4 else 0
scn.set_precision("tf32")
coords, feats, _ = make_batch(list(range(5)), 50)
x = scn.InputLayer(3, 4096, mode=4)([coords, feats.cuda()])
md = x.metadata
size = 4096 >> lvl
level = md.levels[size]
conv = scn.SubmanifoldConvolution(3, cin, cout, 3, False).cuda()
f = torch.randn(level.n, cin, device='cuda')
gw = ops.GemmWeight(conv.weight.view(27, cin, cout))
scn.set_option("halo_dbg", dbg)
for _ in range(3): ops.subm_conv(f, level, gw)
buf = torch.zeros(4096, dtype=torch.int64, device='cuda')
tile = (level.n // 128) // 2
fn = _lib.lib.b200scn_debug_timeline; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert fn(buf.data_ptr(), tile) == 0
ops.subm_conv(f, level, gw); torch.cuda.synchronize()
fn(None, -1)
t = buf.cpu().numpy(); t0 = t[0]
def r(i): return int(t[i] - t0) if t[i] else None
print("dbg", dbg, "level", lvl, "n", level.n, "cin", cin, "cout", cout, "tile", tile)
print("prologue done", r(1), " accum seen", r(2), " epilogue done", r(3))
for kb in range((cin + 31) // 32): print("halo kb", kb, "start", r(32 + 2 * kb), "done", r(33 + 2 * kb))
for g in range(4):
    rows = [(r(64 + g * 256 + 2 * u), r(65 + g * 256 + 2 * u)) for u in range(128) if t[64 + g * 256 + 2 * u]]
    if rows: print("builder group", g, " (slot free, built):", rows[:30])
for g in range(3):
    rows = [(r(64 + g * 256 + 2 * u), r(2200 + g * 512 + 4 * u), r(2201 + g * 512 + 4 * u), r(65 + g * 256 + 2 * u)) for u in range(6) if t[2200 + g * 512 + 4 * u]]
    if rows: print("builder", g, "(slot free, stores issued, stores done, arrived):", rows)
mm = [(r(1088 + 4 * i), r(1089 + 4 * i), r(1090 + 4 * i), r(1091 + 4 * i)) for i in range(64) if t[1088 + 4 * i]]
print("mma visits (wait begins, operands seen, mmas issued, committed):", mm[:40])
print("mma (after mmas, after commit1):", [(r(1600 + 2 * i), r(1601 + 2 * i)) for i in range(128) if t[1600 + 2 * i]][:60])

"""Per-shape kernel timing of one config (CUDA events around every C-ABI conv/BN call)."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/3d-weakly-supervised-semantic-segmentation_b200')
import torch
import sparseconvnet as scn
from sparseconvnet import ops
from b200scn_synth import CONFIGS, build_encoder, make_batch
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3_unet_m32_r2_res_s50_b5"
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32"
kind, m, reps, res, scale, batch = CONFIGS[cfg]
scn.set_precision(prec)
net = build_encoder(scn, kind, m, reps, res).cuda()
coords, feats, _ = make_batch(list(range(batch)), scale)
coords = coords.cuda(); feats = feats.cuda()
def step():
    f = feats.detach().requires_grad_(True)
    y = net([coords, f]); y.mean().backward()
for i in range(3): step()
ops.profile_reserve(2000); ops.profile_begin(); step(); prof = ops.profile_end()
for kind, d in prof.items():
    print("%-12s n=%3d  %.2f ms  %.0f GB/s  %.1f TF/s" % (kind, d["n"], d["ms"], d["bytes"]/d["ms"]/1e6, d["flops"]/d["ms"]/1e9))
    for sh, (ms, n, by, fl) in sorted(d["shapes"].items(), key=lambda kv: -kv[1][0])[:12]:
        cc = int(sh.split('x')[0]) // 2
        print("      CinCout=%-6d %-28s n=%2d  %.3f ms each  %.0f GB/s  %.1f TF/s" % (cc, sh, n, ms/n, by/ms/1e6, fl/ms/1e9))

"""One forward+backward of a config (for ncu captures of individual kernels)."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/3d-weakly-supervised-semantic-segmentation_b200')
import torch
import sparseconvnet as scn
from b200scn_synth import CONFIGS, build_encoder, make_batch
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3_unet_m32_r2_res_s50_b5"
kind, m, reps, res, scale, batch = CONFIGS[cfg]
scn.set_precision(sys.argv[2] if len(sys.argv) > 2 else "tf32")
net = build_encoder(scn, kind, m, reps, res).cuda()
coords, feats, _ = make_batch(list(range(batch)), scale)
f = feats.cuda().requires_grad_(True)
y = net([coords, f]); y.mean().backward()
torch.cuda.synchronize()
print("ok", y.shape)

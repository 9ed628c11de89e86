"""One small tiled SubM layer (fwd, bwd-input, weight gradient) + BN + strided conv for compute-sanitizer:

    compute-sanitizer --tool {memcheck,racecheck,synccheck,initcheck} python tools/sanitize_tiled.py

Small on purpose (the sanitizer serialises and instruments every access): 2 x 12 k points at 2 cm -> ~20 k sites,
~160 tiles, Cin/Cout 32 and 64 (both template instances of halo_conv_tc_kernel), tiled kernel forced on."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "3d-weakly-supervised-semantic-segmentation_b200"))
os.environ["B200SCN_HALO"] = "1"
import torch
import sparseconvnet as scn
from b200scn_synth import make_batch

scn.set_precision("tf32")
coords, feats, _ = make_batch([0, 1], 50, n_points=12000)
net = scn.Sequential(scn.InputLayer(3, 4096, mode=4), scn.SubmanifoldConvolution(3, 3, 32, 3, False),
                     scn.BatchNormReLU(32), scn.SubmanifoldConvolution(3, 32, 32, 3, False),
                     scn.BatchNormReLU(32), scn.SubmanifoldConvolution(3, 32, 64, 3, False),
                     scn.BatchNormReLU(64), scn.SubmanifoldConvolution(3, 64, 64, 3, False),
                     scn.Convolution(3, 64, 96, 2, 2, False), scn.BatchNormReLU(96),
                     scn.SubmanifoldConvolution(3, 96, 96, 3, False),
                     scn.Deconvolution(3, 96, 64, 2, 2, False), scn.OutputLayer(3)).cuda()
f = feats.cuda().requires_grad_(True)
y = net([coords, f])
y.mean().backward()
torch.cuda.synchronize()
print("sanitize_tiled: ok, out", tuple(y.shape), "launches", scn.launch_count(), "checksum %.6f" % float(y.double().sum()))

"""One SubmanifoldConvolution forward + backward (input and weight gradients) at a chosen level of cfg3 (kernel traces)."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/3d-weakly-supervised-semantic-segmentation_b200')
import torch
import sparseconvnet as scn
from b200scn_synth import make_batch
level, C = int(sys.argv[1]), int(sys.argv[2])
scn.set_precision("tf32")
coords, feats, _ = make_batch(list(range(5)), 50)
x = scn.InputLayer(3, 4096, mode=4)([coords, feats.cuda()])
md = x.metadata
size = 4096 >> level
lvl = md.levels[size]
f = torch.randn(lvl.n, C, device='cuda', requires_grad=True)
t = scn.SparseConvNetTensor(f, md, torch.LongTensor([size] * 3))
conv = scn.SubmanifoldConvolution(3, C, C, 3, False).cuda()
for _ in range(3):
    y = conv(t)
    y.features.sum().backward()
torch.cuda.synchronize()
print("level", level, "n", lvl.n, "C", C)

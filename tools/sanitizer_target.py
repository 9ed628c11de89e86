"""Smallest run that launches the shipped tensor-memory kernels once each, for compute-sanitizer (racecheck / synccheck /
memcheck): tiled SubM forward (both template instances), pair-list weight gradient, tile-stationary weight gradient,
gather kernel, BatchNorm forward/backward.   compute-sanitizer --tool racecheck python tools/sanitizer_target.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "3d-weakly-supervised-semantic-segmentation_b200"))
import torch
import sparseconvnet as scn
from sparseconvnet import ops
from b200scn_synth import make_batch

scn.set_precision("tf32")
coords, feats, _ = make_batch([0], 20, n_points=30000)          # ~20 k voxels: above the tiled kernel's threshold
x = scn.InputLayer(3, 4096, mode=4)([coords, feats.cuda()])
lvl = x.metadata.levels[4096]
print("voxels", lvl.n, "tiled:", ops._use_tiled(lvl.n))
torch.manual_seed(0)
for cin, cout in ((32, 32), (64, 128)):
    f = torch.randn(lvl.n, cin, device="cuda")
    g = torch.randn(lvl.n, cout, device="cuda")
    w = torch.randn(27, cin, cout, device="cuda") * 0.1
    y = ops.subm_conv(f, lvl, ops.GemmWeight(w))
    ref = ops.gather_conv(f, lvl.subm_map(), lvl.n, 27, ops.GemmWeight(w), rules=lvl)
    pin, pout, offs = lvl.subm_pairs_ordered(lvl.tile_plan(ops._halo["hcap"]).perm)
    d0 = ops.pair_dw(f, g, pin, pout, offs, 27, lvl.n)
    d1 = ops.subm_dw_tiled(f, g, lvl)
    torch.cuda.synchronize()
    print(cin, cout, "tiled vs gather", float((y - ref).norm() / ref.norm()), "dW tile vs pairs",
          float((d1 - d0).norm() / d0.norm()) if d1 is not None else None)
bn = scn.BatchNormReLU(32).cuda()
t = scn.SparseConvNetTensor(torch.randn(lvl.n, 32, device="cuda", requires_grad=True), x.metadata, x.spatial_size)
out = bn(t)
out.features.sum().backward()
torch.cuda.synchronize()
print("sanitizer target done")

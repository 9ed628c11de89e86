"""Host-side cost of one step: wall time to ENQUEUE fwd+bwd (no sync) and a cProfile of where it goes."""
import cProfile, pstats, sys, time, io
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/3d-weakly-supervised-semantic-segmentation_b200')
import torch
import sparseconvnet as scn
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
    t0 = time.perf_counter()
    y = net([coords, f])
    t1 = time.perf_counter()
    y.mean().backward()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2
for i in range(3): step()
for i in range(3):
    a, b, c = step()
    print("enqueue fwd %.1f ms  bwd %.1f ms  drain %.1f ms" % (a * 1e3, b * 1e3, c * 1e3))
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35); print(s.getvalue()[:6000])

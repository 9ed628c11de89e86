"""Step time of any BASELINE config (train fwd+bwd, or eval forward with val_reps)."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/3d-weakly-supervised-semantic-segmentation_b200')
import torch
import sparseconvnet as scn
from b200scn_synth import CONFIGS, build_encoder, make_batch
cfg, mode = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "train")
kind, m, reps, res, scale, batch = CONFIGS[cfg]
scn.set_precision("tf32")
net = build_encoder(scn, kind, m, reps, res).cuda()
data = [make_batch(list(range(batch)), scale, step=s) for s in range(3)]
data = [(c.cuda(), f.cuda()) for c, f, _ in data]
def step(i):
    c, f = data[i % 3]
    if mode == "train":
        f = f.detach().requires_grad_(True)
        y = net([c, f]); y.mean().backward()
    else:
        with torch.no_grad():
            y = net([c, f])
    return y
if mode != "train": net.eval()
for i in range(4): y = step(i)
torch.cuda.synchronize(); t0 = time.perf_counter()
n = 6
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n): y = step(i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
vox = None
print("%s %s: %.2f ms/step, out %s, peak mem %.1f GB" % (cfg, mode, ms, tuple(y.shape), torch.cuda.max_memory_allocated() / 1e9))

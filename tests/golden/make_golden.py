"""Generate the golden fixtures of tests/golden/ from the CPU oracle (run once, commit the .npz files).

    python tests/golden/make_golden.py

The reference holds no golden vectors for this path, and its sparse-conv dependency cannot be imported
here, so these vectors come from the oracle restatement (oracle/scn_oracle.py), which is itself pinned by literal cases,
dense conv3d equivalence and fp64 gradient checks (tests/test_oracle_*.py).  They freeze the oracle against drift and
give the GPU tests a fixed target that does not depend on the oracle code at run time.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "3d-weakly-supervised-semantic-segmentation_b200"))

from oracle import scn_oracle as ref  # noqa: E402
from _util import random_cloud  # noqa: E402


def small_unet(ns, smooth=False):
    """A 3-level SparseConvUNet (models/SparseConvNet.py:59-71 with nPlanes [8,16,24], residual blocks)."""
    leak = 1.0 if smooth else 0
    net = ns.Sequential(
        ns.InputLayer(3, 4096, mode=4),
        ns.SubmanifoldConvolution(3, 3, 8, 3, False),
        ns.UNet(3, 1, [8, 16, 24], True, leakiness=leak),
        ns.BatchNormLeakyReLU(8, leakiness=leak),
        ns.OutputLayer(3))
    return net


def main():
    # ---- rulebooks
    coords, feats = random_cloud(123, 1200, 14, 2)
    pv, vox = ref.input_rules(coords.numpy())
    nbr = ref.subm_map(vox)
    parent, off, voxc = ref.strided(vox, 2)
    nbr1 = ref.subm_map(voxc)
    np.savez_compressed(os.path.join(HERE, "rulebooks.npz"), coords=coords.numpy(), pv=pv, vox=vox, nbr=nbr,
                        parent=parent, off=off, voxc=voxc, nbr1=nbr1)
    # ---- a small encoder, smooth (leakiness 1) so that gradients are continuous and comparable at 1e-3
    torch.manual_seed(7)
    net = small_unet(ref, smooth=True)
    f = feats.clone().requires_grad_(True)
    out = net([coords, f])
    torch.manual_seed(8)
    go = torch.randn_like(out) / out.shape[0]
    out.backward(go)
    sd = {k: v.detach().numpy() for k, v in net.state_dict().items()}
    grads = {"grad::" + n: p.grad.numpy() for n, p in net.named_parameters()}
    np.savez_compressed(os.path.join(HERE, "small_unet.npz"), coords=coords.numpy(), feats=feats.numpy(),
                        grad_out=go.numpy(), logits=out.detach().numpy(), grad_feats=f.grad.numpy(),
                        **{"param::" + k: v for k, v in sd.items()}, **grads)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()

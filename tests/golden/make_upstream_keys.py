"""Writes tests/golden/upstream_unet_keys.json: the state_dict key names and shapes an upstream sparseconvnet 0.2
checkpoint of the reference's SparseConvUNet (m=16, block_reps=1, VGG blocks; models/SparseConvNet.py:59-71) carries.
sparseconvnet is not installable here (no network, needs sparsehash), so the list is derived from the module tree the
reference itself restates at Function_test.py:113-226 (scn.UNet as nested scn.Sequential / ConcatTable / JoinTable,
children indexed "0", "1", ...) as built by the CPU oracle, with upstream's naming of the BatchNorm buffers
(`runningMean`, `runningVar`: SURVEY.md section 5 / appendix B, [UPSTREAM-RECALL]).   python tests/golden/make_upstream_keys.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-weakly-supervised-semantic-segmentation_b200"))
from b200scn_synth import build_encoder   # noqa: E402
from oracle import scn_oracle as ref      # noqa: E402

net = build_encoder(ref, "SparseConvUNet", 16, 1, False)
keys = []
for k, v in net.state_dict().items():
    k = k.replace("running_mean", "runningMean").replace("running_var", "runningVar")
    keys.append([k, list(v.shape)])
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "upstream_unet_keys.json")
json.dump({"model": "SparseConvUNet m=16 block_reps=1 residual_blocks=False", "keys": keys}, open(out, "w"), indent=0)
print(out, len(keys))

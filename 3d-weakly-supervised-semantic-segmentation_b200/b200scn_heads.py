"""Scene-level heads of the reference, fused on the B200 path.

`MultiLabelHead` mirrors `models/MultiLabelContrastive.py:50-70` (`MultiLabel`): a point-cloud encoder followed by
`nn.Linear(embed_width, NUM_CLASSES)`; state_dict keys are the reference's (`pc_encoder.*`, `linear.weight`,
`linear.bias`), so a reference checkpoint loads.  In training the reference materialises the per-point tensor
`(sum P, C)` (OutputLayer), averages it per scene in a Python loop (`models/SparseConvNet.py:20-26`) and runs the Linear
and `F.multilabel_soft_margin_loss` (`utils/loss.py:21-30`) as ~12 eager kernels.  Here the training path is:
encoder trunk -> `scn.SceneMeanPooling` (OutputLayer + per-scene mean in one pass over the voxel features, the per-point
tensor never exists) -> ONE kernel for Linear + loss (`b200scn_head_multilabel`), one for their backward.
Evaluation keeps the reference's per-point logits (`OutputLayer` + Linear over every point).
"""
import torch
from torch import nn

import sparseconvnet as scn
from sparseconvnet import _lib
from sparseconvnet._lib import check, lib, ptr

NUM_CLASSES = 20   # dataset/data.py: ScanNet benchmark classes


class MultiLabelHeadFn(torch.autograd.Function):
    """(pooled (B,C), weight (NC,C), bias (NC) | None, labels (B,NC) | None) -> logits (B,NC), loss ()"""

    @staticmethod
    def forward(ctx, pooled, weight, bias, labels):
        pooled, weight = pooled.contiguous(), weight.contiguous()
        B, C = pooled.shape
        NC = weight.shape[0]
        dev = pooled.device
        logits = torch.empty((B, NC), dtype=torch.float32, device=dev)
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        if labels is not None:
            labels = labels.to(device=dev, dtype=torch.float32).contiguous()
            assert labels.shape == (B, NC), "labels must be (B, NUM_CLASSES) multi-hot (utils/loss.py:28)"
        scratch = torch.zeros(B + 1, dtype=torch.float32, device=dev)
        check(lib.b200scn_head_multilabel(ptr(pooled), ptr(weight), ptr(bias), ptr(labels), B, C, NC, ptr(logits), ptr(loss),
                                          ptr(scratch), _lib.stream_for(pooled)))
        ctx.save_for_backward(pooled, weight, labels, logits)
        ctx.has_bias = bias is not None
        return logits, loss

    @staticmethod
    def backward(ctx, d_logits, d_loss):
        pooled, weight, labels, logits = ctx.saved_tensors
        B, C = pooled.shape
        NC = weight.shape[0]
        d_pooled = torch.empty_like(pooled)
        d_w = torch.empty_like(weight)
        d_b = torch.empty(NC, dtype=torch.float32, device=pooled.device) if ctx.has_bias else None
        d_logits = d_logits.contiguous() if d_logits is not None else None
        d_loss = d_loss.contiguous() if (d_loss is not None and labels is not None) else None
        check(lib.b200scn_head_multilabel_bwd(ptr(pooled), ptr(weight), ptr(labels), ptr(logits), ptr(d_loss), ptr(d_logits),
                                              B, C, NC, ptr(d_pooled), ptr(d_w), ptr(d_b), _lib.stream_for(pooled)))
        return d_pooled, d_w, d_b, None


def _trunk_and_output(encoder):
    """An encoder Sequential as models/SparseConvNet.py builds it: (..., scn.OutputLayer) -> (modules before it, it)."""
    mods = list(encoder)
    if not isinstance(mods[-1], scn.OutputLayer):
        raise ValueError("the encoder must end with scn.OutputLayer (models/SparseConvNet.py:70,87,157)")
    return mods[:-1], mods[-1]


class MultiLabelHead(nn.Module):
    """Drop-in for the reference's `MultiLabel` model (`models/MultiLabelContrastive.py:50-70`).

    pc_encoder: a reference encoder object (`SparseConvBase_` subclass: has `.encoder`) or the scn.Sequential itself.
    forward(x, istrain=False, labels=None):
      istrain: x = (batch, ...) as train.py:69 passes it, batch = {coords, feature, batch_offsets} (attribute or key access);
               -> (global_logits (B, NUM_CLASSES), None) like the reference, or (global_logits, loss) when `labels` is given;
      else   : x = batch -> per-point logits (sum P, NUM_CLASSES)."""

    def __init__(self, pc_encoder, embed_width, num_classes=NUM_CLASSES):
        super().__init__()
        self.pc_encoder = pc_encoder
        self.linear = nn.Linear(embed_width, num_classes)
        self.pool = scn.SceneMeanPooling()

    def _sequential(self):
        return getattr(self.pc_encoder, "encoder", self.pc_encoder)

    @staticmethod
    def _get(batch, name):
        return batch[name] if isinstance(batch, dict) else getattr(batch, name)

    def head_loss(self, y, batch_size, labels=None):
        """Trunk output (level-0 SparseConvNetTensor) -> (global_logits, loss | None): pooling + Linear (+ loss)."""
        pooled = self.pool(y, batch_size)
        logits, loss = MultiLabelHeadFn.apply(pooled, self.linear.weight, self.linear.bias, labels)
        return logits, (loss if labels is not None else None)

    def forward(self, x, istrain=False, labels=None):
        if istrain:
            x = x[0]
        coords, feats = self._get(x, "coords"), self._get(x, "feature")
        trunk, out_layer = _trunk_and_output(self._sequential())
        y = trunk[0]([coords, feats])
        for mod in trunk[1:]:
            y = mod(y)
        if not istrain:
            return self.linear(out_layer(y))
        return self.head_loss(y, len(self._get(x, "batch_offsets")) - 1, labels)

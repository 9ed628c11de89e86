"""The reference's per-step data path on the GPU (dataset/data.py:165-200 trainMerge, :266-290 valMerge).

Scenes stay resident on the device as fp32 xyz + rgb (12 + 12 bytes per point instead of the 32-byte int64 coordinate row
the reference's collate builds on the host); each step the host draws the same few random numbers per scene the reference
draws (rotation / flip / jitter matrix, offsets) and `merge()` runs the transform, the [0, full_scale) crop, the
truncation to integers and the key packing on the device (csrc/augment.cu), handing `InputLayer` a `PackedKeys` object
instead of a (sum P, 4) LongTensor.  Coordinates are bit-identical to the numpy reference path on the same draws
(tests/test_gpu_datapath.py).
"""
import numpy as np
import torch

from sparseconvnet import _lib
from sparseconvnet._lib import check, lib, ptr
from sparseconvnet.metadata import PackedKeys


def draw_val_params(rng, n_scenes, scale, full_scale=4096):
    """The random draws of valMerge (dataset/data.py:268-277) for n_scenes scenes, in the reference's order, from a numpy
    Generator / RandomState-like `rng` with integers / random / uniform: -> mats (B,3,3), pre0, pre (B,3), r1, r2 (B,3)."""
    mats, pre, r1, r2 = [], [], [], []
    for _ in range(n_scenes):
        m = np.eye(3)
        m[0][0] *= rng.integers(0, 2) * 2 - 1
        m *= scale
        th = rng.random() * 2 * np.pi
        m = np.matmul(m, [[np.cos(th), np.sin(th), 0], [-np.sin(th), np.cos(th), 0], [0, 0, 1]])
        mats.append(m)
        pre.append(rng.uniform(-2, 2, 3))
        r1.append(rng.random(3))
        r2.append(rng.random(3))
    return np.stack(mats), full_scale / 2, np.stack(pre), np.stack(r1), np.stack(r2)


def draw_train_params(rng, n_scenes, scale):
    """The random draws of trainMerge (dataset/data.py:165-177): jittered, flipped, scaled, rotated matrix; no pre-offset."""
    mats, r1, r2 = [], [], []
    for _ in range(n_scenes):
        m = np.eye(3) + rng.standard_normal((3, 3)) * 0.1
        m[0][0] *= rng.integers(0, 2) * 2 - 1
        m *= scale
        th = rng.random() * 2 * np.pi
        mats.append(np.matmul(m, [[np.cos(th), np.sin(th), 0], [-np.sin(th), np.cos(th), 0], [0, 0, 1]]))
        r1.append(rng.random(3))
        r2.append(rng.random(3))
    return np.stack(mats), 0.0, None, np.stack(r1), np.stack(r2)


def merge_numpy(scenes, mats, pre0, pre, r1, r2, form, full_scale=4096):
    """The reference's host path on the same draws (numpy float64, dataset/data.py:172-186 / 272-284), for parity tests and
    as the documentation of what the kernels compute: -> coords (sum P', 4) int64 [x, y, z, b], kept point indices,
    batch_offsets."""
    locs, kept, offs, base = [], [], [0], 0
    for b, (xyz, _rgb) in enumerate(scenes):
        a = np.matmul(xyz, mats[b])                     # float32 points promoted by the float64 matrix
        if pre is not None:
            a = a + pre0 + pre[b]
        m, M = a.min(0), a.max(0)
        if form == 0:
            length = M - m
            offset = -m + np.clip(full_scale - length - 0.001, 0, None) * r1[b] + np.clip(full_scale - length + 0.001, None, 0) * r2[b]
        else:
            offset = -m + np.clip(full_scale - M + m - 0.001, 0, None) * r1[b] + np.clip(full_scale - M + m + 0.001, None, 0) * r2[b]
        a = a + offset
        idxs = (a.min(1) >= 0) * (a.max(1) < full_scale)
        a = torch.from_numpy(a[idxs]).long()
        locs.append(torch.cat([a, torch.full((a.shape[0], 1), b, dtype=torch.long)], 1))
        kept.append(np.nonzero(idxs)[0] + base)
        base += xyz.shape[0]
        offs.append(offs[-1] + int(idxs.sum()))
    return torch.cat(locs, 0), np.concatenate(kept), offs


class DeviceScenes:
    """A batch of scenes resident on one GPU: xyz (sum n, 3) fp32, rgb (sum n, C) fp32, scene_start (B + 1) int32."""

    def __init__(self, scenes, device):
        self.device = torch.device(device)
        self.B = len(scenes)
        starts = np.cumsum([0] + [s[0].shape[0] for s in scenes]).astype(np.int32)
        self.P = int(starts[-1])
        self.scene_start = torch.from_numpy(starts).to(self.device)
        self.xyz = torch.from_numpy(np.ascontiguousarray(np.concatenate([s[0] for s in scenes], 0), dtype=np.float32)).to(self.device)
        self.rgb = torch.from_numpy(np.ascontiguousarray(np.concatenate([s[1] for s in scenes], 0), dtype=np.float32)).to(self.device)
        self.h2d_bytes = self.xyz.numel() * 4 + self.rgb.numel() * 4 + starts.nbytes
        nb = lib.b200scn_augment_scratch_bytes(self.P, self.B)
        self._scratch = torch.empty(nb, dtype=torch.uint8, device=self.device)
        self._nb = nb

    def merge(self, mats, pre0, pre, r1, r2, form, full_scale=4096, jitter=None):
        """One augmentation of the resident scenes -> (PackedKeys, feats (P', C) fp32, batch_offsets list, offsets (B,3) f64
        device tensor).  mats (B,3,3), pre (B,3) or None, r1 / r2 (B,3): float64 numpy, drawn as the reference draws them;
        jitter (B, C) float32 numpy or None: the per-scene colour offset of data.py:200."""
        dev, B, P = self.device, self.B, self.P
        f64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev, non_blocking=True)
        d_m, d_r1, d_r2 = f64(mats), f64(r1), f64(r2)
        d_pre = f64(pre) if pre is not None else None
        keys = torch.empty(max(P, 1), dtype=torch.int64, device=dev)
        kept = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
        hdr = torch.zeros(1 + B, dtype=torch.int32, device=dev)          # [n_kept, kept per scene ...]
        offs = torch.empty((B, 3), dtype=torch.float64, device=dev)
        st = _lib.stream_for(self.xyz)
        check(lib.b200scn_augment_voxelize(ptr(self.xyz), P, ptr(self.scene_start), B, ptr(d_m), float(pre0), ptr(d_pre),
                                           ptr(d_r1), ptr(d_r2), int(form), int(full_scale), ptr(keys), ptr(kept),
                                           hdr.data_ptr(), hdr.data_ptr() + 4, ptr(offs), ptr(self._scratch), self._nb, st))
        C = self.rgb.shape[1]
        d_j = torch.from_numpy(np.ascontiguousarray(jitter, dtype=np.float32)).to(dev, non_blocking=True) if jitter is not None else None
        feats_cap = torch.empty((max(P, 1), C), dtype=torch.float32, device=dev)
        check(lib.b200scn_gather_rows(ptr(self.rgb), C, ptr(kept), hdr.data_ptr(), P, C, ptr(d_j), ptr(self.scene_start), B,
                                      ptr(feats_cap), C, st))
        host = hdr.cpu()      # the caller needs batch_offsets on the host anyway (models/SparseConvNet.py:20-26)
        n_kept = int(host[0])
        batch_offsets = [0]
        for v in host[1:].tolist():
            batch_offsets.append(batch_offsets[-1] + int(v))
        return PackedKeys(keys[:n_kept], batch_size=B), feats_cap[:n_kept], batch_offsets, offs, kept[:n_kept]

"""Shared helpers for the parity tests: random sparse inputs, oracle <-> product comparisons."""
import numpy as np
import torch


def random_cloud(seed, n_points, extent, batch, dup_frac=0.3, full=4096):
    """Random integer coordinates in a small box (so neighbours and duplicates are common), `batch` samples
    concatenated in sample order (as the reference's collate does, dataset/data.py:198)."""
    rng = np.random.default_rng(seed)
    out = []
    for b in range(batch):
        base = rng.integers(0, full - extent, 3)
        c = base + rng.integers(0, extent, (n_points, 3))
        ndup = int(n_points * dup_frac)
        if ndup:
            c[rng.integers(0, n_points, ndup)] = c[rng.integers(0, n_points, ndup)]
        out.append(np.concatenate([c, np.full((n_points, 1), b)], 1))
    coords = torch.from_numpy(np.concatenate(out, 0)).long()
    feats = torch.from_numpy(rng.standard_normal((coords.shape[0], 3)).astype(np.float32))
    return coords, feats


def keys_to_vox(ukeys):
    k = ukeys.cpu().numpy().astype(np.uint64)
    return np.stack([(k >> np.uint64(32)) & np.uint64(0xFFFF), (k >> np.uint64(16)) & np.uint64(0xFFFF),
                     k & np.uint64(0xFFFF), (k >> np.uint64(48)) & np.uint64(0xFFFF)], 1).astype(np.int32)


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def copy_params(src, dst):
    """Copy parameters/buffers from one module tree to another with identical structure."""
    sd = {k: v.detach().clone() for k, v in src.state_dict().items()}
    dst.load_state_dict(sd)


def to_tf32(t):
    """Round an fp32 tensor to the nearest TF32 (10-bit mantissa, ties away from zero: cvt.rna.tf32.f32), so that kernels
    that round their operands and kernels that let the tensor core truncate them see identical, exactly representable inputs."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)

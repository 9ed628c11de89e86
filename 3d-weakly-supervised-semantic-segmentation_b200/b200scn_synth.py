"""Deterministic synthetic ScanNet-shaped scenes and the encoder configurations of BASELINE.json.

Scene: a room (floor + 4 walls) with 12 axis-aligned boxes standing on the floor, points sampled
area-uniformly on the faces with 5 mm noise, centred at the mean (mirrors dataset/ScanNet/prepare_data.py:29-30);
colours U(-1,1).  Batch transform as `valMerge` (dataset/data.py:266-290): rotate about z, multiply by `scale`,
shift into [0, full_scale)^3, drop outliers, truncate to int64, append the sample index.
Pure numpy/torch on the host; used by tests, bench.py and __graft_entry__.smoke().
"""
import numpy as np
import torch


def _rect(origin, u, v):
    return (np.asarray(origin, np.float64), np.asarray(u, np.float64), np.asarray(v, np.float64))


def make_scene(seed, n_points=150000):
    """-> (xyz (n,3) float64 metres, centred; rgb (n,3) float32 in [-1,1])."""
    rng = np.random.default_rng(seed)
    L, W, H = rng.uniform(4, 9), rng.uniform(3, 7), rng.uniform(2.4, 3.0)
    rects = [
        _rect([0, 0, 0], [L, 0, 0], [0, W, 0]),           # floor
        _rect([0, 0, 0], [L, 0, 0], [0, 0, H]), _rect([0, W, 0], [L, 0, 0], [0, 0, H]),
        _rect([0, 0, 0], [0, W, 0], [0, 0, H]), _rect([L, 0, 0], [0, W, 0], [0, 0, H]),
    ]
    for _ in range(12):
        sx, sy, sz = rng.uniform([0.4, 0.4, 0.4], [2.0, 1.2, 1.6])
        ox, oy = rng.uniform(0, max(L - sx, 0.1)), rng.uniform(0, max(W - sy, 0.1))
        rects += [
            _rect([ox, oy, sz], [sx, 0, 0], [0, sy, 0]),  # top
            _rect([ox, oy, 0], [sx, 0, 0], [0, 0, sz]), _rect([ox, oy + sy, 0], [sx, 0, 0], [0, 0, sz]),
            _rect([ox, oy, 0], [0, sy, 0], [0, 0, sz]), _rect([ox + sx, oy, 0], [0, sy, 0], [0, 0, sz]),
        ]
    areas = np.array([np.linalg.norm(np.cross(u, v)) for _, u, v in rects])
    which = rng.choice(len(rects), size=n_points, p=areas / areas.sum())
    a, b = rng.random(n_points), rng.random(n_points)
    O = np.stack([r[0] for r in rects])[which]
    U = np.stack([r[1] for r in rects])[which]
    V = np.stack([r[2] for r in rects])[which]
    xyz = O + a[:, None] * U + b[:, None] * V + rng.normal(0, 0.005, (n_points, 3))
    xyz -= xyz.mean(0)
    rgb = rng.uniform(-1, 1, (n_points, 3)).astype(np.float32)
    return xyz, rgb


def make_batch(scene_seeds, scale, full_scale=4096, n_points=150000, step=0):
    """-> coords (sumP,4) int64 CPU [x,y,z,b], feats (sumP,3) float32 CPU, batch_offsets list.
    `step` re-randomises rotation/offset like a fresh augmentation (fresh coordinates every step)."""
    locs, feats, offs = [], [], [0]
    for bi, seed in enumerate(scene_seeds):
        xyz, rgb = make_scene(seed, n_points)
        rng = np.random.default_rng((seed + 1) * 1000003 + step)
        m = np.eye(3)
        m[0][0] *= rng.integers(0, 2) * 2 - 1
        m *= scale
        th = rng.random() * 2 * np.pi
        m = m @ np.array([[np.cos(th), np.sin(th), 0], [-np.sin(th), np.cos(th), 0], [0, 0, 1]])
        a = xyz @ m + full_scale / 2 + rng.uniform(-2, 2, 3)
        lo, hi = a.min(0), a.max(0)
        a += -lo + np.clip(full_scale - hi + lo - 0.001, 0, None) * rng.random(3) + \
            np.clip(full_scale - hi + lo + 0.001, None, 0) * rng.random(3)
        keep = (a.min(1) >= 0) & (a.max(1) < full_scale)
        a = torch.from_numpy(a[keep]).long()
        locs.append(torch.cat([a, torch.full((a.shape[0], 1), bi, dtype=torch.long)], 1))
        feats.append(torch.from_numpy(rgb[keep]))
        offs.append(offs[-1] + int(keep.sum()))
    return torch.cat(locs, 0), torch.cat(feats, 0), offs


def build_encoder(scn, kind, m, block_reps, residual_blocks, full_scale=4096, dimension=3):
    """The encoders of models/SparseConvNet.py:57-88 composed from the module namespace `scn`
    (this package's `sparseconvnet` or, in tests and the CPU baseline, the oracle)."""
    if kind == "SparseConvUNet":       # models/SparseConvNet.py:59-71
        return scn.Sequential(
            scn.InputLayer(dimension, full_scale, mode=4),
            scn.SubmanifoldConvolution(dimension, 3, m, 3, False),
            scn.UNet(dimension, block_reps, [m, 2 * m, 3 * m, 4 * m, 5 * m, 6 * m, 7 * m], residual_blocks),
            scn.BatchNormReLU(m),
            scn.OutputLayer(dimension))
    if kind == "SparseConvFCNet":      # models/SparseConvNet.py:75-88
        depth = 7
        return scn.Sequential(
            scn.InputLayer(dimension, full_scale, mode=4),
            scn.SubmanifoldConvolution(dimension, 3, m, 3, False),
            scn.FullyConvolutionalNet(dimension, block_reps, [(i + 1) * m for i in range(depth)], residual_blocks,
                                      downsample=[2, 2]),
            scn.BatchNormReLU(depth * (depth + 1) * m // 2),
            scn.OutputLayer(dimension))
    if kind == "SparseConvFCNetDirectUpPool":   # models/SparseConvNet.py:107-158 (embed_length 256)
        nPlanes, ds = [m, 64, 128, 192, 256], [2, 2]

        def block(seq, a, b):
            if residual_blocks:
                seq.add(scn.ConcatTable()
                        .add(scn.Identity() if a == b else scn.NetworkInNetwork(a, b, False))
                        .add(scn.Sequential()
                             .add(scn.BatchNormReLU(a))
                             .add(scn.SubmanifoldConvolution(dimension, a, b, 3, False))
                             .add(scn.BatchNormReLU(b))
                             .add(scn.SubmanifoldConvolution(dimension, b, b, 3, False)))).add(scn.AddTable())
            else:
                seq.add(scn.Sequential()
                        .add(scn.BatchNormReLU(a))
                        .add(scn.SubmanifoldConvolution(dimension, a, b, 3, False)))

        def U(planes):
            seq = scn.Sequential()
            for _ in range(block_reps):
                block(seq, planes[0], planes[0])
            if len(planes) > 1:
                seq.add(scn.Sequential()
                        .add(scn.BatchNormReLU(planes[0]))
                        .add(scn.Convolution(dimension, planes[0], planes[1], ds[0], ds[1], False))
                        .add(U(planes[1:]))
                        .add(scn.UnPooling(dimension, ds[0], ds[1])))
            return seq
        return scn.Sequential(
            scn.InputLayer(dimension, full_scale, mode=4),
            scn.SubmanifoldConvolution(dimension, 3, m, 3, False),
            U(nPlanes),
            scn.BatchNormReLU(nPlanes[-1]),
            scn.OutputLayer(dimension))
    raise ValueError(kind)


# embed width the heads are built with (models/SparseConvNet.py:57,73,107: `embed_length` of the registry entries)
EMBED_WIDTH = {"SparseConvUNet": lambda m: m, "SparseConvFCNet": lambda m: 28 * m, "SparseConvFCNetDirectUpPool": lambda m: 256}

# BASELINE.json configs -> (encoder kind, m, block_reps, residual, scale, batch)
CONFIGS = {
    "cfg1_unet_m16_r1_s20_b1": ("SparseConvUNet", 16, 1, False, 20, 1),
    "cfg2_fcnet_m16_r1_s20_b8": ("SparseConvFCNet", 16, 1, False, 20, 8),
    "cfg3_unet_m32_r2_res_s50_b5": ("SparseConvUNet", 32, 2, True, 50, 5),
    # config/3DUNetWithText_scannet_subcloud_uppool_4gpu.yaml: global batch 30 over 4 GPUs -> 8 scenes per GPU (7.5 rounded up),
    # model MultiLabel = encoder + scene pooling + Linear(256, 20) + multilabel soft margin loss (bench.py uses the fused head)
    "cfg4_uppool_m16_r2_res_s50_b8_head": ("SparseConvFCNetDirectUpPool", 16, 2, True, 50, 8),
    "cfg5_fcnet_m16_r2_res_s100_b6": ("SparseConvFCNet", 16, 2, True, 100, 6),
}

# eval-mode workloads: number of forward passes per step (val_reps; BASELINE.json configs[4] says 3)
EVAL_REPS = {"cfg5_fcnet_m16_r2_res_s100_b6": 3}

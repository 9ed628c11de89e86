"""f2: the reference's per-step data path (dataset/data.py:165-200 trainMerge, :266-290 valMerge) on the GPU, ending in packed
keys (csrc/augment.cu, b200scn_data.DeviceScenes).  Same random draws -> coordinates BIT-IDENTICAL to the numpy float64 host
path, including the [0, full_scale) crop (compaction, batch offsets), and the same voxel numbering out of InputLayer."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scenes(n, npts):
    from b200scn_synth import make_scene
    out = []
    for s in range(n):
        xyz, rgb = make_scene(s, npts + 1000 * s)
        out.append((xyz.astype(np.float32), rgb))
    return out


def _decode(keys):
    k = keys.cpu()
    return torch.stack([(k >> 32) & 0xFFFF, (k >> 16) & 0xFFFF, k & 0xFFFF, (k >> 48) & 0xFFFF], 1)


@pytest.mark.parametrize("form,scale,full,seed", [(1, 50, 4096, 0), (1, 100, 4096, 1), (0, 50, 4096, 2), (0, 20, 4096, 3),
                                                   (1, 50, 256, 4), (0, 50, 300, 5)])
def test_device_merge_matches_the_numpy_path_bit_for_bit(form, scale, full, seed):
    import b200scn_data as D
    scenes = _scenes(3, 30000)
    rng = np.random.default_rng(seed)
    mats, pre0, pre, r1, r2 = (D.draw_val_params(rng, 3, scale, full) if form == 1 else D.draw_train_params(rng, 3, scale))
    want, kept_want, offs_want = D.merge_numpy(scenes, mats, pre0, pre, r1, r2, form, full)
    dev = D.DeviceScenes(scenes, "cuda")
    jitter = rng.standard_normal((3, 3)).astype(np.float32) * 0.1
    pk, feats, offs, offsets, kept = dev.merge(mats, pre0, pre, r1, r2, form, full, jitter=jitter)
    assert offs == offs_want
    assert np.array_equal(kept.cpu().numpy(), kept_want)
    got = _decode(pk.keys)
    assert got.shape == want.shape
    assert torch.equal(got, want), "coordinates differ in %d of %d rows" % (int((got != want).any(1).sum()), want.shape[0])
    if full < 4096:
        assert want.shape[0] < sum(s[0].shape[0] for s in scenes)          # the crop really dropped points
    # features of the kept points + the per-scene jitter (data.py:200)
    rgb = np.concatenate([s[1] for s in scenes], 0)
    scene = np.searchsorted(np.cumsum([s[0].shape[0] for s in scenes]), kept_want, side="right")
    assert np.array_equal(feats.cpu().numpy(), rgb[kept_want] + jitter[scene])


def test_packed_keys_into_input_layer():
    """InputLayer fed with PackedKeys numbers the voxels exactly as when fed with the (sum P, 4) LongTensor."""
    import sparseconvnet as scn
    import b200scn_data as D
    scenes = _scenes(2, 40000)
    rng = np.random.default_rng(7)
    mats, pre0, pre, r1, r2 = D.draw_val_params(rng, 2, 50)
    coords, _, _ = D.merge_numpy(scenes, mats, pre0, pre, r1, r2, 1)
    dev = D.DeviceScenes(scenes, "cuda")
    pk, feats, offs, _, _ = dev.merge(mats, pre0, pre, r1, r2, 1)
    xa = scn.InputLayer(3, 4096, mode=4)([pk, feats])
    xb = scn.InputLayer(3, 4096, mode=4)([coords, feats])
    la, lb = xa.metadata.levels[4096], xb.metadata.levels[4096]
    assert la.n == lb.n and torch.equal(la.ukeys, lb.ukeys) and torch.equal(xa.metadata.pv, xb.metadata.pv)
    assert torch.allclose(xa.features, xb.features, atol=1e-6)      # (means of duplicates: fp32 sums in atomic order)
    assert torch.equal(xa.metadata.levels[64].ukeys, xb.metadata.levels[64].ukeys)     # the whole pyramid

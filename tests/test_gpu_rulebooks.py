"""GPU rulebooks vs the CPU oracle: bit-exact (SURVEY 8c "bit-exact definition")."""
import numpy as np
import pytest
import torch

from _util import keys_to_vox, random_cloud

pytestmark = pytest.mark.gpu


def _build(coords, feats, size=4096, mode=4):
    import sparseconvnet as scn
    inp = scn.InputLayer(3, size, mode=mode)
    return inp([coords, feats.cuda()])


@pytest.mark.parametrize("seed,n,extent,batch", [(0, 500, 12, 2), (1, 4000, 40, 3), (2, 20000, 64, 1), (3, 1, 4, 1)])
def test_input_layer_ids_and_pyramid(seed, n, extent, batch):
    from oracle import scn_oracle as ref
    coords, feats = random_cloud(seed, n, extent, batch)
    x = _build(coords, feats)
    md = x.metadata
    pv_ref, vox_ref = ref.input_rules(coords.numpy())
    assert md.levels[4096].n == vox_ref.shape[0]
    assert np.array_equal(md.pv.cpu().numpy(), pv_ref)
    assert np.array_equal(keys_to_vox(md.levels[4096].ukeys), vox_ref)
    # per-site statistics
    cnt = np.bincount(pv_ref, minlength=vox_ref.shape[0])
    assert np.array_equal(md.count.cpu().numpy(), cnt)
    # strided pyramid: ids, parents, offsets, child maps all equal under the canonical first-touch order
    vox = vox_ref
    size = 4096
    assert md.syncs == 1
    for _ in range(6):
        parent, off, voxc = ref.strided(vox, 2)
        d = md.get_down(size, 2)
        assert d.coarse.n == voxc.shape[0]
        assert np.array_equal(d.parent.cpu().numpy(), parent)
        assert np.array_equal(d.off.cpu().numpy().astype(np.int32), off)
        assert np.array_equal(keys_to_vox(d.coarse.ukeys), voxc)
        child = np.full((voxc.shape[0], 8), -1, np.int32)
        child[parent, off] = np.arange(vox.shape[0], dtype=np.int32)
        assert np.array_equal(d.child_map().cpu().numpy(), child)
        vox, size = voxc, size // 2
    assert md.syncs == 1  # the whole pyramid came from the single InputLayer sync


@pytest.mark.parametrize("seed,n,extent,batch", [(0, 500, 12, 2), (5, 6000, 30, 2)])
def test_subm_rulebook(seed, n, extent, batch):
    from oracle import scn_oracle as ref
    coords, feats = random_cloud(seed, n, extent, batch)
    x = _build(coords, feats)
    md = x.metadata
    _, vox = ref.input_rules(coords.numpy())
    size = 4096
    for lvl in range(3):
        level = md.levels[size]
        nbr_ref = ref.subm_map(vox)
        assert np.array_equal(level.subm_map().cpu().numpy(), nbr_ref)
        assert level.rule_counts() == [int((nbr_ref[:, k] >= 0).sum()) for k in range(27)]
        # scn-form pair lists, canonical order (ascending out inside each offset)
        pin, pout, offs = level.subm_pairs()
        pin, pout, offs = pin.cpu().numpy(), pout.cpu().numpy(), offs.cpu().numpy()
        rules = ref.rules_from_map(nbr_ref)
        for k, (ri, ro) in enumerate(rules):
            assert offs[k + 1] - offs[k] == len(ri)
            assert np.array_equal(pin[offs[k]:offs[k + 1]], ri)
            assert np.array_equal(pout[offs[k]:offs[k + 1]], ro)
        _, _, vox = ref.strided(vox, 2)
        md.get_down(size, 2)
        size //= 2


def test_stride4_on_demand():
    from oracle import scn_oracle as ref
    coords, feats = random_cloud(7, 3000, 50, 2)
    x = _build(coords, feats)
    md = x.metadata
    _, vox = ref.input_rules(coords.numpy())
    parent, off, voxc = ref.strided(vox, 4)
    d = md.get_down(4096, 4)
    assert np.array_equal(d.parent.cpu().numpy(), parent)
    assert np.array_equal(d.off.cpu().numpy().astype(np.int32), off)
    assert np.array_equal(keys_to_vox(d.coarse.ukeys), voxc)
    assert d.coarse.size == 1024 and d.K == 64


def test_input_modes_and_errors():
    import sparseconvnet as scn
    from oracle import scn_oracle as ref
    coords, feats = random_cloud(11, 800, 6, 2, dup_frac=0.6)
    for mode in (1, 2, 3, 4):
        x = _build(coords, feats, mode=mode)
        y = ref.InputLayer(3, 4096, mode=mode)([coords, feats])
        assert torch.allclose(x.features.cpu(), y.features, rtol=1e-5, atol=1e-6), mode
        o = scn.OutputLayer(3)(x).cpu()
        o_ref = ref.OutputLayer(3)(y)
        assert torch.allclose(o, o_ref, rtol=1e-5, atol=1e-6), mode
    bad = coords.clone()
    bad[3, 1] = 4096
    with pytest.raises(ValueError):
        _build(bad, feats)
    with pytest.raises(RuntimeError):
        scn.InputLayer(3, 4096, mode=4)([coords, feats])  # CPU features: no CPU fallback


def test_empty_input():
    import sparseconvnet as scn
    coords = torch.zeros((0, 4), dtype=torch.long)
    feats = torch.zeros((0, 3))
    x = scn.InputLayer(3, 4096, mode=4)([coords, feats.cuda()])
    assert x.features.shape == (0, 3)
    y = scn.SubmanifoldConvolution(3, 3, 8, 3, False).cuda()(x)
    assert y.features.shape == (0, 8)
    assert scn.OutputLayer(3)(y).shape == (0, 8)

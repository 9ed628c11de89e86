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


def _morton(vox):
    def spread(v):
        v = v.astype(np.uint64)
        out = np.zeros_like(v)
        for bit in range(16):
            out |= ((v >> np.uint64(bit)) & np.uint64(1)) << np.uint64(3 * bit)
        return out
    return (vox[:, 3].astype(np.uint64) << np.uint64(48)) | (spread(vox[:, 0]) << np.uint64(2)) | \
        (spread(vox[:, 1]) << np.uint64(1)) | spread(vox[:, 2])


@pytest.mark.parametrize("seed,n,extent,batch,hcap", [(0, 500, 12, 2, 384), (5, 6000, 30, 2, 384), (6, 5000, 20, 3, 32)])
def test_tile_plan(seed, n, extent, batch, hcap):
    """Spatially tiled convolution plan (conv_halo.cu): Morton order bit-exact against numpy; every lmap entry resolves,
    through the tile's halo list, to exactly the neighbour id of the rulebook (or is flagged absent / beyond capacity)."""
    from oracle import scn_oracle as ref
    coords, feats = random_cloud(seed, n, extent, batch)
    x = _build(coords, feats)
    level = x.metadata.levels[4096]
    _, vox = ref.input_rules(coords.numpy())
    nbr = ref.subm_map(vox)
    plan = level.tile_plan(hcap)
    perm = plan.perm.cpu().numpy()
    assert np.array_equal(perm, np.argsort(_morton(vox), kind="stable").astype(np.int32))
    N = vox.shape[0]
    T = (N + 127) // 128
    lmap = plan.lmap.cpu().numpy().view(np.uint16).reshape(T, 27, 128)
    hids = plan.halo_ids.cpu().numpy().reshape(T, hcap)
    hn = plan.halo_n.cpu().numpy()
    km = plan.kmask.cpu().numpy().view(np.uint32)
    saw_overflow = False
    for t in range(T):
        rows = perm[t * 128:(t + 1) * 128]
        sub = nbr[rows]                                   # (r, 27)
        distinct = np.unique(sub[sub >= 0])
        assert hn[t] == min(len(distinct), hcap)
        assert len(np.unique(hids[t, :hn[t]])) == hn[t] and np.isin(hids[t, :hn[t]], distinct).all()
        lm = lmap[t, :, :len(rows)].T                     # (r, 27)
        assert (lmap[t, :, len(rows):] == 0xFFFF).all()
        assert np.array_equal(lm == 0xFFFF, sub < 0)
        inhalo = lm < 0xFFFE
        assert np.array_equal(hids[t][lm[inhalo]], sub[inhalo])
        over = lm == 0xFFFE
        saw_overflow |= bool(over.any())
        assert not np.isin(sub[over], hids[t, :hn[t]]).any()
        assert km[t] == sum(1 << k for k in range(27) if (sub[:, k] >= 0).any())
    assert saw_overflow == (hcap == 32)

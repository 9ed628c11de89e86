"""Parity at BASELINE.json's full size (cfg3: 5 scenes x 150 k points at 2 cm, ~650 k voxels), where the CPU oracle would
take minutes: size-independent properties of the rulebooks and plans, and agreement between independently written
kernels of the same operator (SURVEY 8c property list; the small-size tests compare against the oracle itself)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene():
    import sparseconvnet as scn
    from b200scn_synth import make_batch
    coords, feats, offs = make_batch(list(range(5)), 50)
    x = scn.InputLayer(3, 4096, mode=4)([coords, feats.cuda()])
    return coords, feats, x


def test_voxel_ids_first_occurrence(scene):
    coords, feats, x = scene
    md = x.metadata
    c = coords.numpy()
    key = (c[:, 3] << 48) | (c[:, 0] << 32) | (c[:, 1] << 16) | c[:, 2]
    uniq, first = np.unique(key, return_index=True)
    order = np.argsort(first)                      # ids = rank of first occurrence in input row order
    lvl = md.levels[4096]
    assert lvl.n == uniq.shape[0]
    assert np.array_equal(lvl.ukeys.cpu().numpy().astype(np.int64), uniq[order])
    pv = md.pv.cpu().numpy()
    assert np.array_equal(uniq[order][pv], key)    # every point maps to the voxel that holds its key
    assert np.array_equal(np.bincount(pv, minlength=lvl.n), md.count.cpu().numpy())


def test_rulebook_properties(scene):
    _, _, x = scene
    md = x.metadata
    size = 4096
    for _ in range(3):
        lvl = md.levels[size]
        nbr = lvl.subm_map().long()
        n = lvl.n
        ar = torch.arange(n, device=nbr.device)
        assert torch.equal(nbr[:, 13], ar)                                   # rules[13] is the identity
        for k in range(13):                                                  # rules[k] and rules[26-k] are transposes
            o = ar[nbr[:, k] >= 0]
            i = nbr[o, k]
            assert torch.equal(nbr[i, 26 - k], o)
            assert int((nbr[:, k] >= 0).sum()) == int((nbr[:, 26 - k] >= 0).sum())
        assert lvl.rule_counts() == [int(v) for v in (nbr >= 0).sum(0).tolist()]
        # the canonical pair lists are the map read column by column, ascending out
        pin, pout, offs = lvl.subm_pairs()
        offs = offs.cpu().tolist()
        for k in (0, 13, 26):
            sel = nbr[:, k] >= 0
            assert torch.equal(pout[offs[k]:offs[k + 1]].long(), ar[sel])
            assert torch.equal(pin[offs[k]:offs[k + 1]].long(), nbr[sel, k])
        # strided relation: every fine site has exactly one parent and sits in its parent's child map
        d = md.get_down(size, 2)
        child = d.child_map().long()
        fine = torch.arange(d.fine.n, device=child.device)
        assert torch.equal(child[d.parent.long(), d.off.long()], fine)
        assert int((child >= 0).sum()) == d.fine.n
        size //= 2


def test_tile_plan_invariants(scene):
    _, _, x = scene
    lvl = x.metadata.levels[4096]
    hcap = 384
    plan = lvl.tile_plan(hcap)
    n = lvl.n
    perm = plan.perm.long()
    assert torch.equal(torch.sort(perm)[0], torch.arange(n, device=perm.device))     # a permutation
    T = (n + 127) // 128
    nbr = lvl.subm_map()
    pad = T * 128 - n
    rows = torch.cat([perm, perm.new_zeros(pad)])
    want = nbr[rows].view(T, 128, 27).permute(0, 2, 1).long()                          # (T,27,128) neighbour ids
    if pad:
        want[-1, :, 128 - pad:] = -1
    lmap = (plan.lmap.view(T, 27, 128).long() & 0xFFFF)
    hids = plan.halo_ids.view(T, hcap).long()
    assert torch.equal(lmap == 0xFFFF, want < 0)                                       # absent <-> absent
    inhalo = lmap < 0xFFFE
    got = torch.gather(hids, 1, torch.where(inhalo, lmap, torch.zeros_like(lmap)).view(T, -1)).view(T, 27, 128)
    assert torch.equal(got[inhalo], want[inhalo])                                      # slot -> the rulebook's neighbour
    over = lmap == 0xFFFE
    assert float(over.float().mean()) < 0.01                                           # capacity 384 covers almost all
    own = torch.arange(128, device=lmap.device).expand(T, 128)
    centre = lmap[:, 13, :]
    assert torch.equal(centre[centre != 0xFFFF], own[centre != 0xFFFF])                # own rows hold slots 0..127


@pytest.mark.parametrize("level,c", [(0, 32), (1, 64)])
def test_tiled_and_gather_kernels_agree(scene, level, c):
    import sparseconvnet as scn
    from sparseconvnet import ops
    _, _, x = scene
    lvl = x.metadata.levels[4096 >> level]
    scn.set_precision("tf32")
    try:
        torch.manual_seed(level)
        from _util import to_tf32
        f = to_tf32(torch.randn(lvl.n, c, device="cuda"))   # representable inputs: rounding (tiled) == truncation (gather)
        g = to_tf32(torch.randn(lvl.n, c, device="cuda"))
        w = torch.randn(27, c, c, device="cuda") * 0.1
        gw = ops.GemmWeight(w)
        scn.set_tiled("off")
        ref = ops.subm_conv(f, lvl, gw)
        scn.set_tiled("on")
        out = ops.subm_conv(f, lvl, gw)
        assert float((out - ref).norm() / ref.norm()) < 1e-5          # same TF32 products, different summation order
        # linearity (size-independent property of the operator)
        lin = ops.subm_conv(2.0 * f - 0.5 * g, lvl, gw)
        assert float((lin - (2.0 * out - 0.5 * ops.subm_conv(g, lvl, gw))).norm() / lin.norm()) < 2e-3
        # weight gradient: canonical and Morton-ordered pair lists hold the same pairs
        pin, pout, offs = lvl.subm_pairs()
        dw_a = ops.pair_dw(f, g, pin, pout, offs, 27, lvl.n)
        pin2, pout2, offs2 = lvl.subm_pairs_ordered(lvl.tile_plan(384).perm)
        assert torch.equal(offs, offs2)
        dw_b = ops.pair_dw(f, g, pin2, pout2, offs2, 27, lvl.n)
        assert float((dw_a - dw_b).norm() / dw_a.norm()) < 1e-5
    finally:
        scn.set_tiled("auto")
        scn.set_precision("fp32")

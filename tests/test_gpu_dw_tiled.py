"""Tile-stationary weight gradient (csrc/dw_tile.cu, b200scn_subm_dw_tiled) vs the pair-list kernel and the fp32 oracle
formula, on the full-size cfg3 grids; bit-reproducible across runs (fixed-order reduction of the CTA partials)."""
import pytest
import torch

from _util import to_tf32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def levels():
    import sparseconvnet as scn
    from b200scn_synth import make_batch
    coords, feats, _ = make_batch(list(range(5)), 50)
    x = scn.InputLayer(3, 4096, mode=4)([coords, feats.cuda()])
    md = x.metadata
    return [md.levels[4096 >> i] for i in range(4)]


# (level, Ca, Cg): cfg3 shapes the dispatcher sends to the tiled weight gradient + corner cases (Ca % 32 != 0, Cg % 32 != 0)
SHAPES = [(0, 32, 32), (0, 64, 32), (1, 64, 64), (1, 128, 64), (2, 96, 96), (3, 128, 128), (1, 48, 80), (2, 40, 16),
          (1, 8, 112), (3, 64, 256)]


@pytest.mark.parametrize("level,ca,cg", SHAPES)
def test_tiled_dw_matches_pair_kernel_and_is_reproducible(levels, level, ca, cg):
    import sparseconvnet as scn
    from sparseconvnet import ops
    lvl = levels[level]
    scn.set_precision("tf32")
    try:
        torch.manual_seed(1000 * level + ca + cg)
        a = to_tf32(torch.randn(lvl.n, ca, device="cuda"))      # representable inputs: rounding == truncation
        g = to_tf32(torch.randn(lvl.n, cg, device="cuda"))
        pin, pout, offs = lvl.subm_pairs()
        ref = ops.pair_dw(a, g, pin, pout, offs, 27, lvl.n)
        dw = ops.subm_dw_tiled(a, g, lvl)
        assert dw is not None, "shape not taken by the tiled weight gradient"
        assert dw.shape == ref.shape == (27, ca, cg)
        err = float((dw - ref).norm() / ref.norm())
        assert err < 6e-5, err   # measured <= 3.6e-5 (fp32 sums in different orders: 128-row tiles vs 4096-pair chunks)
        # against exact fp64 on one offset (k = 4) and the centre (k = 13)
        nbr = lvl.subm_map().long()
        for k in (4, 13, 26):
            sel = nbr[:, k] >= 0
            want = a[nbr[sel, k]].double().t() @ g[sel].double()
            assert float((dw[k].double() - want).norm() / want.norm()) < 1e-4   # fp32 accumulation over ~1e5 rules
        for _ in range(10):
            assert torch.equal(ops.subm_dw_tiled(a, g, lvl), dw)
    finally:
        scn.set_precision("fp32")


def test_tiled_dw_overflow_slots(levels):
    import sparseconvnet as scn
    from sparseconvnet import ops
    lvl = levels[1]
    scn.set_precision("tf32")
    old = ops._halo["hcap"]
    try:
        ops.set_halo_capacity(64)      # most neighbours beyond the halo capacity: the global-map route
        torch.manual_seed(5)
        a = to_tf32(torch.randn(lvl.n, 64, device="cuda"))
        g = to_tf32(torch.randn(lvl.n, 32, device="cuda"))
        pin, pout, offs = lvl.subm_pairs()
        ref = ops.pair_dw(a, g, pin, pout, offs, 27, lvl.n)
        dw = ops.subm_dw_tiled(a, g, lvl)
        assert float((dw - ref).norm() / ref.norm()) < 1e-5
    finally:
        ops.set_halo_capacity(old)
        lvl.plan = None
        scn.set_precision("fp32")


@pytest.mark.parametrize("level,ca,cg", [(0, 32, 32), (1, 64, 64), (3, 128, 128), (2, 40, 16)])
def test_blocked_pair_table_and_kernel_variants(levels, level, ca, cg):
    """b200scn_pair_lists_blocked: same lists as b200scn_pair_lists_ordered + a row-block table whose segments partition
    every offset's list at the Morton-block boundaries; b200scn_pair_dw_blocked and the 64-pairs-per-stage variant of
    b200scn_pair_dw agree with the default kernel (fp32 sums in another order)."""
    import sparseconvnet as scn
    from sparseconvnet import ops
    from sparseconvnet.metadata import build_pairs, pair_row_block
    lvl = levels[level]
    scn.set_precision("tf32")
    try:
        perm = lvl.tile_plan(ops._halo["hcap"]).perm
        total = sum(lvl.rule_counts())
        pin, pout, offs = build_pairs(lvl.subm_map(), lvl.n, 27, total, order=perm)
        rb = pair_row_block(lvl.n, 27)
        bin_, bout, boffs, (blk, nblk) = build_pairs(lvl.subm_map(), lvl.n, 27, total, order=perm, row_block=rb)
        assert torch.equal(bin_[:total], pin[:total]) and torch.equal(bout[:total], pout[:total]) and torch.equal(boffs, offs)
        blk_h, offs_h = blk.cpu().long(), offs.cpu().long()
        assert nblk == (lvl.n + rb - 1) // rb and blk_h.numel() == 27 * nblk + 1
        assert bool((blk_h[1:] >= blk_h[:-1]).all()) and int(blk_h[-1]) == total
        assert torch.equal(blk_h[0:27 * nblk:nblk], offs_h[:27])          # block 0 of offset k starts where list k starts
        # every pair of segment (k, b) has its row inside Morton block b
        rank = torch.empty(lvl.n, dtype=torch.long, device="cuda")
        rank[perm.long()] = torch.arange(lvl.n, device="cuda")
        seg = torch.bucketize(torch.arange(total), blk_h[1:], right=True)   # flat segment index of every pair
        assert torch.equal((rank[pout[:total].long()].cpu() // rb), seg % nblk)
        torch.manual_seed(7)
        a = to_tf32(torch.randn(lvl.n, ca, device="cuda"))
        g = to_tf32(torch.randn(lvl.n, cg, device="cuda"))
        ref = ops.pair_dw(a, g, pin, pout, offs, 27, lvl.n)
        got = ops.pair_dw_blocked(a, g, pin, pout, blk, nblk, 27)
        assert got is not None and float((got - ref).norm() / ref.norm()) < 3e-5
        scn.set_option("dw_pairs", 64)
        got64 = ops.pair_dw(a, g, pin, pout, offs, 27, lvl.n)
        scn.set_option("dw_pairs", 32)
        assert float((got64 - ref).norm() / ref.norm()) < 3e-5
    finally:
        scn.set_option("dw_pairs", 32)
        scn.set_precision("fp32")

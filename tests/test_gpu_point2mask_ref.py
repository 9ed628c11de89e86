"""A12 parity PINNED BY THE REFERENCE ITSELF: `b200scn_p2m_*` (through the C ABI) against the reference's own point2mask
extension, built unmodified from /root/reference/ops/point2mask/_ext_src by oracle/build_ref.py into oracle/_ref/ (the
prebuilt .so travels to the GPU box; nothing here reads /root/reference).  Indices and gathered values bit-exact; the
gradient (fp32 atomics in the reference, order-dependent) to 1e-5.

Shapes: the reference's own smoke configuration radius=4, nsample=3 (ops/point2mask/point2mask_modules.py:427-429) and the
production blur radius=1, nsample=20 (ops/pseudo_dataset_generator/configs.py:11-12) on a 256 x 256 pixel grid
(65 536 queries, preprocess_mask.py:31-32), with padded instances (`pointsnum`, point2mask_modules.py:213-231)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_ext():
    from oracle import build_ref
    mod = build_ref.load()
    if mod is None:
        pytest.fail("oracle/_ref/point2mask_ext.so missing: run `python oracle/build_ref.py` in the build container")
    return mod


def _case(b, n, res, seed, pad=True):
    rng = np.random.default_rng(seed)
    xy = (rng.random((b, n, 2)) * res).astype(np.float32)
    # pixel-centre queries, as Pixel2Mask builds them (a meshgrid over H x W)
    gx, gy = np.meshgrid(np.arange(res, dtype=np.float32) + 0.5, np.arange(res, dtype=np.float32) + 0.5, indexing="ij")
    q = np.broadcast_to(np.stack([gx.ravel(), gy.ravel()], 1)[None], (b, res * res, 2)).copy()
    ptnum = rng.integers(n // 4, n, b).astype(np.int32) if pad else np.zeros(b, np.int32)
    ptnum[0] = n        # the largest instance: the reference's `k < n - ptnum` bound leaves it nothing to scan
    if b > 1:
        ptnum[1] = 0
    return torch.from_numpy(q).cuda(), torch.from_numpy(xy).cuda(), torch.from_numpy(ptnum).cuda()


@pytest.mark.parametrize("b,n,res,radius,nsample,seed", [
    (4, 20000, 64, 4.0, 3, 0),          # reference smoke parameters
    (3, 30000, 128, 1.0, 20, 1),        # production blur parameters
    (2, 200000, 256, 1.0, 20, 2),       # production shape per instance: 65 536 queries x 200 k points
    (2, 5000, 32, 0.75, 7, 3),
])
def test_against_reference_extension(ref_ext, b, n, res, radius, nsample, seed):
    import point2mask_ext as ext
    q, xy, ptnum = _case(b, n, res, seed)
    idx_ref = ref_ext.ball_query(q, xy, ptnum, radius, nsample)
    idx = ext.ball_query(q, xy, ptnum, radius, nsample, bucketed=False)   # shared-memory scan kernel
    assert idx.dtype == idx_ref.dtype == torch.int32 and idx.shape == idx_ref.shape
    assert torch.equal(idx, idx_ref)                                    # bit-exact incl. the -1 sentinel
    idx_b = ext.ball_query(q, xy, ptnum, radius, nsample, bucketed=True)  # cell-bucketed kernel: identical
    assert torch.equal(idx_b, idx_ref)
    assert int((idx_ref[0] >= 0).sum()) == 0                            # n - ptnum == 0: nothing scanned
    torch.manual_seed(seed)
    feats = torch.randn(b, 2, n, device="cuda")
    g_ref = ref_ext.group_points(feats, idx_ref)
    g = ext.group_points(feats, idx)
    assert torch.equal(g, g_ref)
    go = torch.randn_like(g_ref)
    gp_ref = ref_ext.group_points_grad(go, idx_ref, n)
    gp = ext.group_points_grad(go, idx, n)
    assert gp.shape == gp_ref.shape
    assert float((gp - gp_ref).abs().max()) <= 1e-5 * max(1.0, float(gp_ref.abs().max()))


def test_oracle_restatement_matches_reference(ref_ext):
    """The C restatement used by the CPU-side tests (oracle/scn_rules.c) is pinned to the reference too."""
    from oracle import scn_oracle as ref
    q, xy, ptnum = _case(2, 3000, 32, 5)
    ptnum[0] = 100
    idx_ref = ref_ext.ball_query(q, xy, ptnum, 2.0, 5).cpu().numpy()
    idx_o = ref.ball_query(2.0, 5, xy.cpu().numpy(), q.cpu().numpy(), ptnum.cpu().numpy())
    assert np.array_equal(idx_o, idx_ref)


def test_production_shape_bucketed(ref_ext):
    """ops/pseudo_dataset_generator/configs.py:11-12 + preprocess_mask.py:20,31-32: blur radius 1, 20 samples, 256 x 256
    pixel-centre queries, ~200 k points per instance, padded instances; 8 of the 64 instances of a production batch (the
    reference extension scans O(m n) per instance, so the full batch would take it minutes).  Bit-exact vs the reference,
    and the speed-up is printed for profiles/."""
    import time
    import point2mask_ext as ext
    b, n, res = 8, 200000, 256
    q, xy, ptnum = _case(b, n, res, 11)
    ptnum[:] = torch.randint(0, n // 3, (b,), dtype=torch.int32)
    ptnum[0] = 0
    ptnum = ptnum.cuda()

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        return out, time.perf_counter() - t0

    idx_ref, t_ref = timed(lambda: ref_ext.ball_query(q, xy, ptnum, 1.0, 20))
    idx_scan, t_scan = timed(lambda: ext.ball_query(q, xy, ptnum, 1.0, 20, bucketed=False))
    idx_b, t_b = timed(lambda: ext.ball_query(q, xy, ptnum, 1.0, 20, bucketed=True))
    assert torch.equal(idx_scan, idx_ref) and torch.equal(idx_b, idx_ref)
    pts = float(b) * res * res
    print("point2mask ball_query B=%d m=%d n=%d r=1 nsample=20: reference ext %.1f ms, scan kernel %.1f ms, bucketed %.2f ms "
          "(%.0fx vs reference; %.1f M queries/s)" % (b, res * res, n, 1e3 * t_ref, 1e3 * t_scan, 1e3 * t_b, t_ref / t_b, pts / t_b / 1e6))

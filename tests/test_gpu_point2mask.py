"""ops/point2mask kernels on the GPU vs the literal CPU restatement (oracle/scn_rules.c), bit-exact indices."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("b,n,m,radius,nsample,seed", [(2, 3000, 4096, 4.0, 3, 0), (3, 5000, 1024, 1.0, 20, 1),
                                                        (1, 100, 64, 0.5, 4, 2), (2, 2049, 300, 3.0, 5, 3)])
def test_ball_query_and_grouping(b, n, m, radius, nsample, seed):
    import point2mask_ext as ext
    from oracle import scn_oracle as ref
    rng = np.random.default_rng(seed)
    side = int(np.sqrt(m))
    xy = (rng.random((b, n, 2)) * side).astype(np.float32)
    q = (rng.random((b, m, 2)) * side).astype(np.float32)
    # padded instances: the reference scans only the first n - ptnum candidates (ball_query_gpu.cu:28)
    ptnum = rng.integers(0, n // 2, b).astype(np.int32)
    ptnum[0] = 0
    idx_ref = ref.ball_query(radius, nsample, xy, q, ptnum)
    idx = ext.ball_query(torch.from_numpy(q).cuda(), torch.from_numpy(xy).cuda(), torch.from_numpy(ptnum).cuda(), radius, nsample)
    assert idx.dtype == torch.int32 and np.array_equal(idx.cpu().numpy(), idx_ref)
    feats = rng.standard_normal((b, 2, n)).astype(np.float32)
    g_ref = ref.group_points(feats, idx_ref)
    g = ext.group_points(torch.from_numpy(feats).cuda(), idx)
    assert np.array_equal(g.cpu().numpy(), g_ref)
    go = rng.standard_normal(g_ref.shape).astype(np.float32)
    gp_ref = ref.group_points_grad(go, idx_ref, n)
    gp = ext.group_points_grad(torch.from_numpy(go).cuda(), idx, n)
    assert np.allclose(gp.cpu().numpy(), gp_ref, rtol=1e-5, atol=1e-5)


def test_reference_checks():
    import point2mask_ext as ext
    q = torch.zeros(1, 4, 2).cuda()
    xy = torch.zeros(1, 8, 2).cuda()
    with pytest.raises(RuntimeError):      # the reference's default float `pointsnum` is rejected by CHECK_IS_INT
        ext.ball_query(q, xy, torch.zeros(1).cuda(), 1.0, 2)
    with pytest.raises(RuntimeError):
        ext.ball_query(q.cpu(), xy.cpu(), torch.zeros(1, dtype=torch.int32), 1.0, 2)

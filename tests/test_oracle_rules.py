"""Pin the CPU oracle's integer half with hand-written literal cases (the reference holds no golden vectors for
this path) and with structural properties of scn rulebooks."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import scn_oracle as ref


def test_input_rules_literal():
    # two samples in a 4x4x4 grid, duplicates, first-occurrence ids global over samples
    coords = np.array([
        [1, 1, 1, 0],   # id 0
        [2, 1, 1, 0],   # id 1
        [1, 1, 1, 0],   # dup of 0
        [1, 1, 2, 0],   # id 2
        [1, 1, 1, 1],   # id 3 (same xyz, other sample)
        [2, 1, 1, 0],   # dup of 1
        [0, 0, 0, 1],   # id 4
        [3, 3, 3, 1],   # id 5
        [1, 1, 1, 1],   # dup of 3
    ], np.int64)
    pv, vox = ref.input_rules(coords)
    assert pv.tolist() == [0, 1, 0, 2, 3, 1, 4, 5, 3]
    assert vox.tolist() == [[1, 1, 1, 0], [2, 1, 1, 0], [1, 1, 2, 0], [1, 1, 1, 1], [0, 0, 0, 1], [3, 3, 3, 1]]


def test_subm_map_literal():
    vox = np.array([[1, 1, 1, 0], [2, 1, 1, 0], [1, 1, 2, 0], [1, 1, 1, 1]], np.int32)
    nbr = ref.subm_map(vox)
    exp = np.full((4, 27), -1, np.int32)
    exp[:, 13] = [0, 1, 2, 3]
    # site0 (1,1,1): +x neighbour is site1 -> k = 9*2+3*1+1 = 22 ; +z neighbour is site2 -> k = 9+3+2 = 14
    exp[0, 22] = 1
    exp[0, 14] = 2
    exp[1, 4] = 0      # site1: -x neighbour is site0 -> k = 0+3+1
    exp[2, 12] = 0     # site2: -z neighbour is site0 -> k = 9+3+0
    # site1 (2,1,1) vs site2 (1,1,2): d = (-1,0,+1) -> k = 0+3+2 = 5 ; reverse (+1,0,-1) -> k = 18+3+0 = 21
    exp[1, 5] = 2
    exp[2, 21] = 1
    assert np.array_equal(nbr, exp)   # sample 1's site never pairs with sample 0's


def test_strided_literal():
    vox = np.array([[0, 0, 0, 0], [1, 1, 1, 0], [2, 0, 1, 0], [3, 1, 0, 0], [0, 0, 0, 1], [1, 0, 0, 1]], np.int32)
    parent, off, voxc = ref.strided(vox, 2)
    assert parent.tolist() == [0, 0, 1, 1, 2, 2]
    assert off.tolist() == [0, 7, 1, 6, 0, 4]     # ((x%2)*2 + y%2)*2 + z%2
    assert voxc.tolist() == [[0, 0, 0, 0], [1, 0, 0, 0], [0, 0, 0, 1]]


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 10 ** 6), st.integers(1, 400), st.integers(2, 9))
def test_rulebook_properties(seed, n, extent):
    rng = np.random.default_rng(seed)
    coords = np.concatenate([rng.integers(0, extent, (n, 3)), rng.integers(0, 2, (n, 1))], 1).astype(np.int64)
    pv, vox = ref.input_rules(coords)
    N = vox.shape[0]
    # ids are first-occurrence ranks; voxels are unique
    assert len({tuple(v) for v in vox.tolist()}) == N
    first = {}
    for r, v in enumerate(pv.tolist()):
        first.setdefault(v, r)
    assert [first[v] for v in range(N)] == sorted(first.values())
    assert np.array_equal(vox[pv], coords.astype(np.int32))
    nbr = ref.subm_map(vox)
    # rules[13] is the identity; rules[k] and rules[26-k] are transposes
    assert np.array_equal(nbr[:, 13], np.arange(N))
    for k in range(13):
        o = np.nonzero(nbr[:, k] >= 0)[0]
        i = nbr[o, k]
        assert np.array_equal(nbr[i, 26 - k], o)
    # strided rules partition the fine set; parent coordinates are floor(fine/2); coarse ids first-touch ordered
    parent, off, voxc = ref.strided(vox, 2)
    assert np.array_equal(voxc[parent][:, :3], vox[:, :3] // 2)
    assert np.array_equal(voxc[parent][:, 3], vox[:, 3])
    _, firsts = np.unique(parent, return_index=True)
    assert np.all(np.diff(firsts) > 0)
    assert len({(int(p), int(o)) for p, o in zip(parent, off)}) == N


def test_input_modes_literal():
    coords = torch.tensor([[1, 1, 1, 0], [1, 1, 1, 0], [2, 2, 2, 0], [1, 1, 1, 0]])
    feats = torch.tensor([[1.0], [2.0], [10.0], [6.0]])
    exp = {1: [6.0, 10.0], 2: [1.0, 10.0], 3: [9.0, 10.0], 4: [3.0, 10.0]}   # Function_test.py:38-44
    for mode, e in exp.items():
        x = ref.InputLayer(3, 8, mode=mode)([coords, feats])
        assert x.features[:, 0].tolist() == e, mode
    # OutputLayer copies voxel features back to every row, no division (App. B.3)
    x = ref.InputLayer(3, 8, mode=4)([coords, feats])
    assert ref.OutputLayer(3)(x)[:, 0].tolist() == [3.0, 3.0, 10.0, 3.0]


def test_point2mask_oracle_literal():
    xy = np.array([[[0, 0], [1, 0], [0.5, 0.5], [5, 5], [0.1, 0.1]]], np.float32)
    q = np.array([[[0, 0], [5, 5]]], np.float32)
    # pointnums = 1 => scan bound n - ptnum = 4: the last point is never seen (ball_query_gpu.cu:28)
    idx = ref.ball_query(1.01, 3, xy, q, np.array([1], np.int32))
    assert idx.tolist() == [[[0, 1, 2], [3, -1, -1]]]
    pts = np.arange(10, dtype=np.float32).reshape(1, 2, 5)
    g = ref.group_points(pts, idx)
    assert g[0, 0].tolist() == [[0, 1, 2], [3, 0, 0]] and g[0, 1].tolist() == [[5, 6, 7], [8, 0, 0]]
    gp = ref.group_points_grad(np.ones_like(g), idx, 5)
    assert gp[0, 0].tolist() == [1, 1, 1, 1, 0]

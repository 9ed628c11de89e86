"""The hot path through the C ABI ALONE (include/b200scn.h), bound with ctypes as INTEGRATION.md shows: InputLayer (pack +
hash numbering + mean features) -> stem SubmanifoldConvolution -> Morton ordering + tile plan -> tiled tensor-memory
SubmanifoldConvolution -> BatchNormReLU -> OutputLayer.  torch is used ONLY to own device memory (torch.empty / zeros and one
read-back of the site count) -- no torch operator runs between the library calls, and nothing of the Python package
`sparseconvnet` (modules, ops, metadata, autograd) is involved.  Checked against the CPU oracle."""
import ctypes
import os

import pytest
import torch

from _util import rel_err

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "3d-weakly-supervised-semantic-segmentation_b200", "sparseconvnet", "libb200scn.so")

vp, i64, i32, f32, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_size_t


def _bind():
    lib = ctypes.CDLL(LIB)
    sig = {
        "b200scn_hash_capacity": (i64, [i64]),
        "b200scn_grid_scratch_bytes": (sz, [i64]),
        "b200scn_pack_coords": (i32, [vp, i64, i32, i64, vp, vp, vp]),
        "b200scn_grid_build": (i32, [vp, i64, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
        "b200scn_input_features": (i32, [vp, i64, i32, vp, vp, vp, vp, i32, vp, vp]),
        "b200scn_subm_map": (i32, [vp, i64, vp, vp, vp, i64, i64, vp, vp, vp]),
        "b200scn_gather_conv": (i32, [vp, i64, i64, vp, i64, i32, vp, i32, i32, vp, i64, vp, i64, i32, vp]),
        "b200scn_morton_perm_scratch_bytes": (sz, [i64]),
        "b200scn_morton_perm": (i32, [vp, i64, i64, i32, vp, vp, sz, vp]),
        "b200scn_tile_plan": (i32, [vp, vp, i64, i32, vp, vp, vp, vp, vp]),
        "b200scn_prep_weight_tf32": (i32, [vp, i32, i32, i32, i32, i32, vp, vp]),
        "b200scn_subm_conv_tiled": (i32, [vp, i64, vp, vp, vp, vp, vp, vp, i32, i64, vp, i32, i32, vp, i64, vp, i64, i32, vp]),
        "b200scn_bn_scratch_doubles": (sz, [i32]),
        "b200scn_bn_forward": (i32, [vp, i64, i64, i32, vp, vp, vp, vp, vp, vp, f32, f32, i32, f32, vp, i64, vp, i32, vp]),
        "b200scn_output_features": (i32, [vp, i64, i64, i32, vp, vp, vp, i32, vp, vp]),
        "b200scn_last_error": (ctypes.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def test_hot_path_through_the_c_abi_alone():
    from b200scn_synth import make_batch
    from oracle import scn_oracle as ref
    lib = _bind()

    def ok(rc):
        assert rc == 0, lib.b200scn_last_error().decode()

    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream(dev).cuda_stream
    E = lambda *shape, dtype=torch.float32: torch.empty(shape, dtype=dtype, device=dev)   # memory only
    Z = lambda *shape, dtype=torch.float32: torch.zeros(shape, dtype=dtype, device=dev)
    p = lambda t: t.data_ptr()

    coords, feats, _ = make_batch([0, 1], 50, n_points=40000)     # ~55 k voxels: above the tiled kernel's usual threshold
    P, size, C0, C1 = coords.shape[0], 4096, 32, 32
    torch.manual_seed(0)
    w0 = torch.randn(27, 3, C0) * (2.0 / 3 / 27) ** 0.5
    w1 = torch.randn(27, C0, C1) * (2.0 / C0 / 27) ** 0.5
    gamma, beta = torch.rand(C1) + 0.5, torch.randn(C1) * 0.1

    # ---- device memory
    d_coords, d_feats = coords.to(dev), feats.to(dev)
    d_w0, d_w1, d_gamma, d_beta = w0.to(dev), w1.to(dev), gamma.to(dev), beta.to(dev)
    keys, hdr = E(P, dtype=torch.int64), Z(2, dtype=torch.int32)            # hdr = [err, n0]
    cap = lib.b200scn_hash_capacity(P)
    sbytes = lib.b200scn_grid_scratch_bytes(P)
    hkeys, hvals, scratch = E(cap, dtype=torch.int64), E(cap, dtype=torch.int32), E(sbytes, dtype=torch.uint8)
    pv, ukeys = E(P, dtype=torch.int32), E(P, dtype=torch.int64)
    first, last, count = E(P, dtype=torch.int32), E(P, dtype=torch.int32), E(P, dtype=torch.int32)

    # ---- InputLayer: site numbering
    ok(lib.b200scn_pack_coords(p(d_coords), P, 4, size, p(keys), p(hdr), st))
    ok(lib.b200scn_grid_build(p(keys), P, None, p(hkeys), p(hvals), cap, p(pv), p(ukeys), p(first), p(last), p(count),
                              p(hdr) + 4, p(scratch), sbytes, st))
    err, n0 = [int(v) for v in hdr.cpu()]                                   # the one host read-back (sizes the outputs)
    assert err == 0 and 0 < n0 <= P
    # ---- InputLayer: features (mode 4 = mean)
    x0 = Z(n0, 3)
    ok(lib.b200scn_input_features(p(d_feats), P, 3, p(pv), p(count), p(first), p(last), 4, p(x0), st))
    # ---- rulebook + stem convolution (Cin = 3: fp32 CUDA cores)
    nbr, counts27 = E(n0, 27, dtype=torch.int32), Z(27, dtype=torch.int32)
    ok(lib.b200scn_subm_map(p(ukeys), n0, None, p(hkeys), p(hvals), cap, size, p(nbr), p(counts27), st))
    x1 = E(n0, C0)
    ok(lib.b200scn_gather_conv(p(x0), 3, n0, p(nbr), n0, 27, p(d_w0), 3, C0, None, 0, p(x1), C0, 0, st))
    # ---- tile plan: Morton ordering inside the library, then halo lists
    hcap, T = 384, (n0 + 127) // 128
    mbytes = lib.b200scn_morton_perm_scratch_bytes(n0)
    mscr, perm = E(mbytes, dtype=torch.uint8), E(n0, dtype=torch.int32)
    ok(lib.b200scn_morton_perm(p(ukeys), n0, size, 15, p(perm), p(mscr), mbytes, st))
    lmap, hids = E(T * 27 * 128, dtype=torch.int16), E(T * hcap, dtype=torch.int32)
    hn, kmask = E(T, dtype=torch.int32), E(T, dtype=torch.int32)
    ok(lib.b200scn_tile_plan(p(nbr), p(perm), n0, hcap, p(lmap), p(hids), p(hn), p(kmask), st))
    # ---- tiled tensor-memory convolution (TF32 tcgen05)
    wkm, x2 = E(27, C1, C0), E(n0, C1)
    ok(lib.b200scn_prep_weight_tf32(p(d_w1), 27, C0, C1, 0, 0, p(wkm), st))
    ok(lib.b200scn_subm_conv_tiled(p(x1), C0, p(nbr), p(perm), p(lmap), p(hids), p(hn), p(kmask), hcap, n0, p(wkm), C0, C1,
                                   None, 0, p(x2), C1, 1, st))
    # ---- BatchNormReLU (training statistics)
    rm, rv, sm_, si = Z(C1), Z(C1) + 1, E(C1), E(C1)
    bscr = Z(lib.b200scn_bn_scratch_doubles(C1), dtype=torch.float64)
    x3 = E(n0, C1)
    ok(lib.b200scn_bn_forward(p(x2), C1, n0, C1, p(d_gamma), p(d_beta), p(rm), p(rv), p(sm_), p(si), 1e-4, 0.9, 1, 0.0,
                              p(x3), C1, p(bscr), 0, st))
    # ---- OutputLayer
    out = E(P, C1)
    ok(lib.b200scn_output_features(p(x3), C1, P, C1, p(pv), p(first), p(last), 4, p(out), st))
    torch.cuda.synchronize()

    # ---- the same net on the CPU oracle
    net = ref.Sequential(ref.InputLayer(3, size, mode=4), ref.SubmanifoldConvolution(3, 3, C0, 3, False),
                         ref.SubmanifoldConvolution(3, C0, C1, 3, False), ref.BatchNormReLU(C1), ref.OutputLayer(3))
    with torch.no_grad():
        net[1].weight.copy_(w0.view(27, 1, 3, C0))
        net[2].weight.copy_(w1.view(27, 1, C0, C1))
        net[3].weight.copy_(gamma)
        net[3].bias.copy_(beta)
    want = net([coords, feats])
    assert out.shape == want.shape
    assert rel_err(out, want) < 1e-3
    assert bool((bscr == 0).all())              # the self-cleaning BatchNorm scratch is zero again
    assert rel_err(rm, net[3].running_mean) < 2e-3 and rel_err(rv, net[3].running_var) < 1e-3   # statistics of a TF32 product

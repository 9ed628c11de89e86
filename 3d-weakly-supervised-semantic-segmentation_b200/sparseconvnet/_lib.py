"""ctypes binding of the b200scn C ABI (include/b200scn.h).

The CUDA library is the only implementation: there is no CPU or eager-PyTorch fallback.  If
libb200scn.so is missing the import fails loudly with the build command.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200scn.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "b200scn CUDA library not built: %s is missing. Run "
        "`python 3d-weakly-supervised-semantic-segmentation_b200/build.py` (needs nvcc, targets sm_100a)." % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

_vp, _i64, _i32, _f32, _sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/b200scn.h one to one
SIGNATURES = {
    "b200scn_last_error": (ctypes.c_char_p, []),
    "b200scn_version": (_i32, []),
    "b200scn_set_device": (_i32, [_i32]),
    "b200scn_set_option": (_i32, [ctypes.c_char_p, _i32]),
    "b200scn_launch_count": (ctypes.c_ulonglong, []),
    "b200scn_hash_capacity": (_i64, [_i64]),
    "b200scn_grid_scratch_bytes": (_sz, [_i64]),
    "b200scn_pack_coords": (_i32, [_vp, _i64, _i32, _i64, _vp, _vp, _vp]),
    "b200scn_grid_build": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200scn_coarse_keys": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp, _vp]),
    "b200scn_subm_map": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "b200scn_child_map": (_i32, [_vp, _vp, _i64, _vp, _i32, _vp, _i64, _vp]),
    "b200scn_pair_scratch_bytes": (_sz, [_i64, _i32]),
    "b200scn_pair_lists": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200scn_pair_lists_ordered": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200scn_pair_lists_blocked": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200scn_gather_conv_tf32_ok": (_i32, [_i32, _i32, _i64]),
    "b200scn_gather_conv": (_i32, [_vp, _i64, _i64, _vp, _i64, _i32, _vp, _i32, _i32, _vp, _i64, _vp, _i64, _i32, _vp]),
    "b200scn_morton_keys": (_i32, [_vp, _i64, _vp, _vp]),
    "b200scn_morton_perm_scratch_bytes": (_sz, [_i64]),
    "b200scn_morton_perm": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _sz, _vp]),
    "b200scn_tile_plan": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "b200scn_subm_conv_tiled": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _i32, _i32, _vp, _i64, _vp, _i64, _i32, _vp]),
    "b200scn_subm_dw_tiled_scratch_bytes": (_sz, [_i64, _i32, _i32, _i32]),
    "b200scn_subm_dw_tiled": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _vp, _vp, _sz, _vp]),
    "b200scn_prep_weight_tf32": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "b200scn_prep_weight_tf32_both": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "b200scn_prep_weight_tf32_batch": (_i32, [_vp, _i32, _i64, _vp]),
    "b200scn_scatter_conv": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _i32, _i32, _vp, _i64, _i32, _vp]),
    "b200scn_group_tiles": (_i32, [_vp, _i32, _i64, _vp, _vp]),
    "b200scn_grouped_conv": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _i64, _i32, _vp, _i32, _i32, _vp, _i64, _vp]),
    "b200scn_pair_dw": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _vp, _i32, _vp]),
    "b200scn_pair_dw_blocked": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "b200scn_unpool": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _i64, _vp]),
    "b200scn_unpool_bwd": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _vp, _i64, _vp]),
    "b200scn_augment_scratch_bytes": (_sz, [_i64, _i32]),
    "b200scn_augment_voxelize": (_i32, [_vp, _i64, _vp, _i32, _vp, ctypes.c_double, _vp, _vp, _vp, _i32, _i64, _vp, _vp, _vp, _vp,
                                        _vp, _vp, _sz, _vp]),
    "b200scn_gather_rows": (_i32, [_vp, _i64, _vp, _vp, _i64, _i32, _vp, _vp, _i32, _vp, _i64, _vp]),
    "b200scn_bn_forward": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _i32, _f32, _vp, _i64, _vp, _i32, _vp]),
    "b200scn_bn_backward": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _f32, _i32, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp]),
    "b200scn_bn_scratch_doubles": (_sz, [_i32]),
    "b200scn_input_features": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "b200scn_input_features_bwd": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "b200scn_output_features": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _vp, _i32, _vp, _vp]),
    "b200scn_output_features_bwd": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _i32, _vp, _i64, _vp]),
    "b200scn_site_rows_scratch_bytes": (_sz, [_i64]),
    "b200scn_site_rows": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "b200scn_output_features_bwd_csr": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _vp]),
    "b200scn_scene_mean": (_i32, [_vp, _i64, _vp, _vp, _i32, _i64, _i32, _i32, _vp, _vp, _vp]),
    "b200scn_scene_mean_bwd": (_i32, [_vp, _vp, _vp, _i32, _vp, _i64, _i32, _vp, _i64, _vp]),
    "b200scn_head_multilabel": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "b200scn_head_multilabel_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "b200scn_maxpool": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _vp, _i64, _vp]),
    "b200scn_maxpool_bwd": (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _vp, _i64, _vp]),
    "b200scn_sparse_to_dense": (_i32, [_vp, _i64, _vp, _i64, _i32, _i64, _vp, _vp]),
    "b200scn_sparse_to_dense_bwd": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _i64, _vp]),
    "b200scn_p2m_ball_query": (_i32, [_i32, _i32, _i32, _f32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "b200scn_p2m_ball_query_scratch_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "b200scn_p2m_ball_query_bucketed": (_i32, [_i32, _i32, _i32, _f32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _sz, _vp]),
    "b200scn_p2m_group_points": (_i32, [_i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "b200scn_p2m_group_points_grad": (_i32, [_i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header and library out of sync
    _fn.restype = _res
    _fn.argtypes = _args


class B200SCNError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise B200SCNError(lib.b200scn_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


_bound_device = [None]


def stream_for(t):
    """Bind this library's runtime to t's device and return torch's current stream handle."""
    if not t.is_cuda:
        raise B200SCNError("b200scn needs CUDA tensors (no CPU fallback); got device %s" % t.device)
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if _bound_device[0] != idx:
        check(lib.b200scn_set_device(idx))
        _bound_device[0] = idx
    return torch.cuda.current_stream(idx).cuda_stream


def set_option(name, value):
    """Kernel tuning / test knob of the C library (include/b200scn.h: b200scn_set_option)."""
    check(lib.b200scn_set_option(name.encode(), int(value)))


# environment overrides are read ONCE, here, never on the launch path
for _env, _opt in (("B200SCN_TC_TMA", "tc_tma"), ("B200SCN_TC_MSUB", "tc_msub"), ("B200SCN_TC_NSPLIT", "tc_nsplit"),
                   ("B200SCN_DW_CHUNK", "dw_chunk"), ("B200SCN_HALO_PF", "halo_pf"), ("B200SCN_HALO_CTAS", None)):
    if os.environ.get(_env):
        if _opt is None:
            set_option("halo_one_cta", 1 if os.environ[_env] == "1" else 0)
        else:
            set_option(_opt, int(os.environ[_env]))


def launch_count():
    """Kernels enqueued by the library since load (bench.py's gpu_launches)."""
    return int(lib.b200scn_launch_count())


def round_rows(n):
    """Row capacity for an n-row buffer: n rounded up so that only 4 significant bits survive (<= 12.5 % slack).
    Site counts change a little every step (fresh augmentation); quantised capacities make every per-level buffer
    size repeat exactly from step to step, so torch's caching allocator reuses blocks instead of fragmenting and
    falling back to (synchronising) cudaMalloc."""
    if n <= 4096:
        return (n + 255) // 256 * 256 if n > 0 else 256
    step = 1 << (n.bit_length() - 4)
    return (n + step - 1) // step * step


def alloc_rows(n, C, device, dtype=torch.float32, zero=False):
    """(n, C) tensor that is a prefix view of a quantised-capacity buffer."""
    cap = round_rows(n)
    buf = (torch.zeros if zero else torch.empty)((cap, C), dtype=dtype, device=device)
    return buf[:n]


def alloc_flat(n, device, dtype, zero=False):
    cap = round_rows(n)
    buf = (torch.zeros if zero else torch.empty)(cap, dtype=dtype, device=device)
    return buf[:n]

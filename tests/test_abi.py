"""The C-ABI library loads without a GPU and exports every symbol include/b200scn.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200scn.h")
LIB = os.path.join(ROOT, "3d-weakly-supervised-semantic-segmentation_b200", "sparseconvnet", "libb200scn.so")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200scn_\w+)\s*\(", src)))


def test_header_declares_the_hot_path():
    names = _declared()
    for must in ("b200scn_grid_build", "b200scn_subm_map", "b200scn_gather_conv", "b200scn_scatter_conv",
                 "b200scn_pair_dw", "b200scn_bn_forward", "b200scn_bn_backward", "b200scn_input_features",
                 "b200scn_output_features", "b200scn_unpool", "b200scn_p2m_ball_query", "b200scn_p2m_group_points"):
        assert must in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build it with python 3d-weakly-supervised-semantic-segmentation_b200/build.py"
    lib = ctypes.CDLL(LIB)
    for name in _declared():
        assert hasattr(lib, name), name
    lib.b200scn_version.restype = ctypes.c_int
    assert lib.b200scn_version() >= 1
    lib.b200scn_hash_capacity.restype = ctypes.c_int64
    lib.b200scn_hash_capacity.argtypes = [ctypes.c_int64]
    cap = lib.b200scn_hash_capacity(1000)
    assert cap >= 2000 and cap & (cap - 1) == 0


def test_python_binding_matches_header():
    import sparseconvnet._lib as L
    assert sorted(L.SIGNATURES) == _declared()
    # no torch types in the ABI: every argument is a plain pointer, integer or float
    allowed = {ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_double}
    for name, (res, args) in L.SIGNATURES.items():
        assert set(args) <= allowed, name


def test_no_cpu_fallback_and_loud_failure():
    import pytest
    import torch
    import sparseconvnet as scn
    coords = torch.zeros((4, 4), dtype=torch.long)
    with pytest.raises(RuntimeError):
        scn.InputLayer(3, 4096, mode=4)([coords, torch.zeros(4, 3)])   # CPU features are refused, not emulated
    src = open(os.path.join(ROOT, "3d-weakly-supervised-semantic-segmentation_b200", "sparseconvnet", "ops.py")).read()
    assert "oracle" not in src   # the product never routes through the test oracle


def test_integration_doc_names_every_entry():
    """INTEGRATION.md maps each C entry to the reference interface it replaces (or says what it is for)."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in _declared() if n not in doc]
    assert not missing, missing

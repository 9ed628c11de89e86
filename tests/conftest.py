import os
import sys

import pytest

# tests compare against the fp32 oracle at tight tolerances unless they pick the TF32 path themselves
os.environ.setdefault("B200SCN_PRECISION", "fp32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "3d-weakly-supervised-semantic-segmentation_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    """Build the native pieces if they are missing (the in-tree .so normally travels with the repo)."""
    import importlib.util
    import shutil
    lib = os.path.join(PKG, "sparseconvnet", "libb200scn.so")
    if not os.path.exists(lib) and shutil.which("nvcc") or (not os.path.exists(lib) and os.path.exists("/usr/local/cuda/bin/nvcc")):
        spec = importlib.util.spec_from_file_location("b200scn_build", os.path.join(PKG, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()


_ensure_built()

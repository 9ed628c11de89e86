"""Build recipe for oracle/_ref: the REFERENCE's own point2mask CUDA extension, compiled unmodified.

TEST INFRASTRUCTURE ONLY (tests/test_gpu_point2mask_ref.py compares `b200scn_p2m_*` against it bit for bit).

    python oracle/build_ref.py            # needs /root/reference (only present in the build container)

Sources are compiled where they lie -- /root/reference/ops/point2mask/_ext_src/{src,include} (the files the reference's
own `ops/point2mask/setup.py` globs) -- with torch's cpp_extension (ninja + nvcc + g++), the only change being the
architecture list: the reference hard-codes TORCH_CUDA_ARCH_LIST="3.7+PTX;...;7.5" (setup.py:19), which nvcc 12.9
rejects, so it is built for 10.0 (sm_100).  nvcc cross-compiles without a GPU.  Output: oracle/_ref/point2mask_ext.so
(git-ignored, NOT gpurun-ignored: it travels to the GPU box, where /root/reference does not exist).
No reference source is copied into this repository.

The sparse-convolution half of the path has no buildable reference (sparseconvnet 0.2 is an absent third-party
dependency), so this is the only `_ref` artefact.
"""
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference/ops/point2mask/_ext_src"
NAME = "point2mask_ext"


def lib_path():
    return os.path.join(OUT_DIR, NAME + ".so")


def build(force: bool = False):
    """-> path of the built extension, or None when /root/reference is absent (GPU box: use the prebuilt file)."""
    if not os.path.isdir(REF_SRC):
        return lib_path() if os.path.exists(lib_path()) else None
    srcs = sorted(glob.glob(os.path.join(REF_SRC, "src", "*.cpp")) + glob.glob(os.path.join(REF_SRC, "src", "*.cu")))
    if (not force) and os.path.exists(lib_path()) and \
            os.path.getmtime(lib_path()) >= max(os.path.getmtime(s) for s in srcs):
        return lib_path()
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
    from torch.utils import cpp_extension
    cpp_extension.load(name=NAME, sources=srcs, extra_include_paths=[os.path.join(REF_SRC, "include")],
                       extra_cflags=["-O3"], extra_cuda_cflags=["-O3"], build_directory=OUT_DIR,
                       is_python_module=False, verbose=False)
    # ninja leaves objects and its own files beside the .so; only the extension itself needs to travel
    for f in os.listdir(OUT_DIR):
        if not f.endswith(".so"):
            os.remove(os.path.join(OUT_DIR, f))
    return lib_path()


def load():
    """Import the prebuilt reference extension (GPU tests only)."""
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    p = lib_path()
    if not os.path.exists(p):
        return None
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

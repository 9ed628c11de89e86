"""Build recipe for the CPU oracle's C half (test infrastructure only).

`python oracle/build.py` compiles oracle/scn_rules.c with gcc into oracle/_build/liboracle.so.
There is no compilable reference implementation of the sparse-conv path under /root/reference
(sparseconvnet 0.2 is an absent third-party dependency; the in-repo point2mask sources are
CUDA + ATen only), so no oracle/_ref is produced -- see DESIGN.md "Oracle".
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")
SRC = os.path.join(HERE, "scn_rules.c")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force) and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-o", LIB, SRC]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))

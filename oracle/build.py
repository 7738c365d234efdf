"""Build the C oracle (``oracle/lbm_ref.c``) into ``oracle/_build/liblbm_oracle.so``.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

There is no ``oracle/_ref``: the reference's implementation of this path is a
GLSL fragment shader plus browser JavaScript (no C/C++ sources to compile), and
no JS engine exists in the build image, so the reference itself cannot be run.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "lbm_ref.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "liblbm_oracle.so")

# -ffp-contract=off: no FMA contraction; -fno-fast-math: IEEE semantics.
CFLAGS = ["-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
          "-Wall", "-Wextra"]


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = ["gcc", *CFLAGS, "-o", OUT, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

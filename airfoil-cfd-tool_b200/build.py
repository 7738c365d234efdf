"""Build libaerolab_lbm.so for sm_100a with nvcc (in-tree, no JIT cache).

    python airfoil-cfd-tool_b200/build.py [--force] [--verbose]

Flags that matter for correctness ("Arithmetic contract"):
  -fmad=false                      nvcc must not contract a*b+c into an FMA
  -Xcompiler -ffp-contract=off     same for the host-side float code
  (-prec-div/-prec-sqrt stay at their IEEE defaults; no --use_fast_math)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "aerolab_lbm", "_lib")
OUT = os.path.join(OUT_DIR, "libaerolab_lbm.so")
SOURCES = ["alb_api.cu", "alb_step.cu", "alb_step2.cu", "alb_march.cu", "alb_geometry.cu", "alb_diag.cu", "alb_particles.cu"]
DEPS = SOURCES + ["alb_common.cuh", "alb_lbm.cuh", os.path.join("..", "..", "include", "aerolab_lbm.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2",
    "-shared", "-cudart", "static", "--threads", "0",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    paths = [os.path.join(CSRC, d) for d in DEPS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(p) <= t for p in paths)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = OUT) -> str:
    """`defines`/`out` build a tuning variant (e.g. defines=["ALB_ST_HINT=1"]) next to the default."""
    if not force and out == OUT and up_to_date():
        return OUT
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [nvcc_path(), *NVCC_FLAGS] + [f"-D{d}" for d in defines]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv or bool(defs), verbose="--verbose" in sys.argv, defines=defs,
                out=os.path.abspath(outs[0]) if outs else OUT))

"""CPU: the three-instruction division by the uniform tau used in the step kernels
(alb_lbm.cuh: div_by_tau<DM_FAST3>) equals IEEE-754 division for every operand the kernel can see.

The library never assumes this: whenever tau changes it runs the same exhaustive comparison on the
device (divtau_check_kernel) and falls back to true division for a tau that fails
(tests/test_gpu_div.py).  Here: exhaustive over all fp32 values with magnitude in [2^-40, 2^8) for
the reference's tau = 0.58 and other practical relaxation times, strided for a set of random ones.
(Operands in the kernel are differences of populations: exactly 0 or at least ~2^-32 in magnitude,
and below 4.)"""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def fbits(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


@pytest.fixture(scope="module")
def chk(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("divchk") / "libdivchk.so")
    flags = ["-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared"]
    hw_fma = "fma" in open("/proc/cpuinfo").read().split("flags", 1)[-1].split("\n", 1)[0].split()
    if hw_fma:
        flags.append("-mfma")
    subprocess.run(["gcc", *flags, "-o", out, os.path.join(HERE, "c", "div_check.c"), "-lm"], check=True)
    L = C.CDLL(out)
    L.div_check.restype = C.c_long
    L.div_check.argtypes = [C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
    L.hw_fma = hw_fma
    return L


@pytest.mark.parametrize("tau", [0.58, 0.505, 0.51, 0.52, 0.55, 0.6, 0.62, 0.75, 0.9, 1.0, 1.5, 2.0])
def test_div_by_tau_exhaustive(chk, tau):
    lo, hi = fbits(2.0 ** -40), fbits(2.0 ** 8)
    stride = 1 if chk.hw_fma else 257          # software fmaf is ~50x slower
    bad = C.c_uint32(0)
    n = chk.div_check(np.float32(tau), lo, hi, stride, C.byref(bad))
    assert n == 0, f"tau={tau}: {n} mismatches, e.g. x bits 0x{bad.value:08x}"


def test_div_by_tau_random_taus(chk):
    rng = np.random.default_rng(5)
    lo, hi = fbits(2.0 ** -40), fbits(2.0 ** 8)
    for tau in np.concatenate([rng.uniform(0.5001, 2.0, 24), [0.5 + 2.0 ** -12, 1.9999999]]):
        bad = C.c_uint32(0)
        n = chk.div_check(np.float32(tau), lo, hi, 1021, C.byref(bad))
        assert n == 0, f"tau={tau}: {n} mismatches, e.g. x bits 0x{bad.value:08x}"

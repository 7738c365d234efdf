"""CPU: tiling of the fused two-step kernel (host logic, alb_debug_step2_plan; no device needed).

Every lattice column must belong to exactly one strip's output range, strips must respect the
kernel's alignment rules (output width a multiple of 4 cells = 16 bytes, at most strip width - 8),
and the segments must cover rows 2 .. ny_local-1."""
import ctypes as C

import pytest


def plan(lib, nx, nyl, nsm=148):
    out = (C.c_int * 5)()
    assert lib.alb_debug_step2_plan(nx, nyl, nsm, out) == 0
    return dict(nstrips=out[0], wo=out[1], hs=out[2], ntiles=out[3], wi=out[4])


@pytest.mark.parametrize("nx", [128, 130, 320, 504, 505, 633, 1100, 2048, 4096, 8191, 8192, 32768, 100000])
@pytest.mark.parametrize("nyl", [1, 2, 3, 4, 7, 66, 160, 1024, 2048, 16384, 65000])
def test_plan_covers_the_lattice(built_lib, nx, nyl):
    p = plan(built_lib, nx, nyl)
    pitch = (nx + 127) // 128 * 128
    assert p["wo"] % 4 == 0 and 0 < p["wo"] <= p["wi"] - 8
    assert p["nstrips"] * p["wo"] >= pitch                 # the strips reach the end of the row ...
    assert (p["nstrips"] - 1) * p["wo"] < pitch            # ... and none of them is empty
    rows = nyl - 2
    if rows <= 0:
        assert p["ntiles"] == 0
        return
    nsegs = -(-rows // p["hs"])
    assert p["hs"] >= 1 and p["ntiles"] == p["nstrips"] * nsegs
    assert nsegs * p["hs"] >= rows > (nsegs - 1) * p["hs"]  # segments cover rows 2 .. nyl-1, none empty


def test_plan_prefers_whole_waves(built_lib):
    # 32768 x 16384 on 148 SMs: the chosen segment height must not leave a mostly empty last wave
    p = plan(built_lib, 32768, 16384)
    waves = p["ntiles"] / 148
    assert 64 <= p["hs"] <= 256
    assert waves - int(waves) > 0.6 or waves == int(waves), p
    # a slab of an 8-GPU run
    q = plan(built_lib, 32768, 2048)
    wq = q["ntiles"] / 148
    assert wq - int(wq) > 0.6 or wq == int(wq), q


def test_plan_rejects_nonsense(built_lib):
    out = (C.c_int * 5)()
    assert built_lib.alb_debug_step2_plan(0, 10, 148, out) == -1
    assert built_lib.alb_debug_step2_plan(100, 10, 0, out) == -1
    assert built_lib.alb_debug_step2_plan(100, 10, 148, None) == -1

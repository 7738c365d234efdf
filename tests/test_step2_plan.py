"""CPU: tiling of the fused two-step kernel (host logic, alb_debug_step2_plan; no device needed).

march2_kernel hands (row segment x column segment) units to independent warps.  A warp stages 128
columns and writes the middle 120, so: the column segments must cover every column that can be
deep, no staged column may lie outside the row, and the row segments must cover rows
2 .. ny_local-1.  Widths that are not a multiple of 4, or below 256: the first and the last task
hold the borders and cannot be deep, the segments cover [128, pitch - 128).  Otherwise the inlet
and outlet columns are part of the domain; the first segment stages [0, 128) and writes [0, 124),
the last stages [nx - 128, nx) (16-byte aligned) and writes [nx - 124, nx), the ones between write
120 columns each from column 124 on."""
import ctypes as C

import pytest


def plan(lib, nx, nyl, nsm=148):
    out = (C.c_int * 5)()
    assert lib.alb_debug_step2_plan(nx, nyl, nsm, out) == 0
    return dict(nseg=out[0], wo=out[1], hs=out[2], nunits=out[3], warps=out[4], wi=128)


@pytest.mark.parametrize("nx", [128, 130, 252, 256, 260, 320, 384, 385, 504, 505, 633, 1100, 2000, 2048, 4096, 8191, 8192, 32768, 100000])
@pytest.mark.parametrize("nyl", [1, 2, 3, 4, 7, 66, 160, 1024, 2048, 16384, 65000])
def test_plan_covers_the_lattice(built_lib, nx, nyl):
    p = plan(built_lib, nx, nyl)
    pitch = (nx + 127) // 128 * 128
    assert p["wo"] == 120 and p["warps"] in (12, 16)
    rows = nyl - 2
    if nx % 4 == 0 and nx >= 256:
        nmid = p["nseg"] - 2
        assert nmid >= 0
        assert 124 + nmid * p["wo"] >= nx - 124                 # the middle segments reach the last segment's columns ...
        assert nmid == 0 or 124 + (nmid - 1) * p["wo"] < nx - 124      # ... and none of them is empty
        assert nmid == 0 or 120 * nmid + p["wi"] <= nx          # middle segment s stages [120 s, 120 s + 128)
        if rows > 0:
            nsegs = -(-rows // p["hs"])
            assert p["hs"] >= 1 and p["nunits"] == p["nseg"] * nsegs
            assert nsegs * p["hs"] >= rows > (nsegs - 1) * p["hs"]
        else:
            assert p["nunits"] == 0
        return
    if pitch < 384:
        assert p["nseg"] == 0 and p["nunits"] == 0          # no task can be deep: first and last task hold the borders
        return
    first_staged = 124                                      # outputs of segment s: [128 + 120 s, 248 + 120 s)
    assert 128 + p["nseg"] * p["wo"] >= pitch - 128         # the segments reach the last column that can be deep ...
    assert 128 + (p["nseg"] - 1) * p["wo"] < pitch - 128    # ... and none of them is empty
    assert first_staged + (p["nseg"] - 1) * p["wo"] + p["wi"] <= pitch   # staged columns stay inside the row
    if rows <= 0:
        assert p["nunits"] == 0
        return
    nsegs = -(-rows // p["hs"])
    assert p["hs"] >= 1 and p["nunits"] == p["nseg"] * nsegs
    assert nsegs * p["hs"] >= rows > (nsegs - 1) * p["hs"]  # row segments cover rows 2 .. nyl-1, none empty


@pytest.mark.parametrize("nx,nyl", [(32768, 16384), (32768, 2048), (4096, 2048), (2048, 1024)])
def test_plan_balances_the_warps(built_lib, nx, nyl):
    # units per warp slot (148 SMs x warps per CTA): the pass lasts ceil(units / warps) units, so the
    # chosen height must keep the overhead of the last partial round and of the two extra step-1 rows small
    p = plan(built_lib, nx, nyl)
    warps = 148 * p["warps"]
    rounds = -(-p["nunits"] // warps)
    ideal = nx / 120 * (nyl - 2) * 2 / warps                # row-steps per warp if the work split perfectly
    model = rounds * (2 * p["hs"] + 2)
    assert model <= 1.25 * ideal + 8, (p, model, ideal)


def test_plan_rejects_nonsense(built_lib):
    out = (C.c_int * 5)()
    assert built_lib.alb_debug_step2_plan(0, 10, 148, out) == -1
    assert built_lib.alb_debug_step2_plan(100, 10, 0, out) == -1
    assert built_lib.alb_debug_step2_plan(100, 10, 148, None) == -1

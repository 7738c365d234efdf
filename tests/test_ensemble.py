"""Ensemble alpha sweeps: case assignment and CSV on CPU, a small polar on the GPU."""
import csv

import pytest


def test_round_robin_assignment():
    from aerolab_lbm.ensemble import DEFAULT_ALPHAS, assign_cases
    assert len(DEFAULT_ALPHAS) == 31 and DEFAULT_ALPHAS[0] == -10.0 and DEFAULT_ALPHAS[-1] == 20.0
    per_rank = [assign_cases(31, 8, r) for r in range(8)]
    assert [len(p) for p in per_rank] == [4, 4, 4, 4, 4, 4, 4, 3]
    assert sorted(i for p in per_rank for i in p) == list(range(31))


def test_polar_csv(tmp_path):
    from aerolab_lbm.ensemble import write_polar_csv
    rows = [dict(alpha=2.0, CL=0.31, CD=0.052, **{"L/D": 5.96}, CL_me=0.3, CD_me=0.06, sep_frac=0.0,
                 Status="Attached", Re=2504.0, steps=100)]
    p = tmp_path / "polar.csv"
    write_polar_csv(rows, str(p))
    got = list(csv.reader(open(p)))
    assert got[0][:5] == ["α (°)", "CL", "CD", "L/D", "Status"] and got[1][:5] == ["2.0", "0.3100", "0.05200", "6.0", "Attached"]


@pytest.mark.gpu
def test_small_sweep_on_gpu(built_lib):
    import aerolab_lbm as al
    from aerolab_lbm.ensemble import alpha_sweep
    from oracle import geometry as ogeo
    from oracle import lbm as olbm
    rows = alpha_sweep(al.SHAPES["naca0012"](), [-4.0, 0.0, 4.0, 8.0], nx=320, ny=160, steps=1200,
                       settle_steps=600, me_window=256)
    assert [r["alpha"] for r in rows] == [-4.0, 0.0, 4.0, 8.0]
    cl = [r["CL"] for r in rows]
    assert cl[0] < cl[1] < cl[2] < cl[3]                # lift grows with alpha
    assert abs(cl[1]) < 0.05 * abs(cl[3])               # symmetric section at alpha = 0
    assert cl[0] == pytest.approx(-cl[2], rel=0.15)     # near-antisymmetry (staircase mask is not exact)
    # one case against the oracle driven the same way
    o = olbm.OracleTunnel(320, 160)
    o.apply_geometry(ogeo.SHAPES["naca0012"](), 8.0)
    o.step(600)
    for _ in range(50):
        o.step(12)
        o.compute_forces()
    assert rows[3]["CL"] == pytest.approx(o.cl_smooth, rel=1e-12)
    assert rows[3]["CD"] == pytest.approx(o.cd_smooth, rel=1e-12)
    assert rows[3]["Status"] == o.stall_state()[0]

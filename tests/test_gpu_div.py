"""The fused kernels compute ux = jx/rho, uy = jy/rho with a shared reciprocal (div_pair in
alb_step2.cu) instead of two generic divisions.  Inside its accepted operand range it must equal
IEEE division bit for bit; outside it must decline (the kernels then divide for real)."""
import pytest

pytestmark = pytest.mark.gpu


def test_shared_reciprocal_division_matches_ieee(built_lib):
    import aerolab_lbm as al
    with al.WindTunnel(64, 32, 0) as t:
        total = {"checked": 0, "accepted": 0, "wrong": 0}
        for seed in (1, 2, 3, 4):
            r = t.selftest_division(pairs=3 << 30, seed=seed)      # two quotients per operand triple
            for k in total:
                total[k] += r[k]
    assert total["wrong"] == 0, total
    assert total["checked"] >= 4 * (3 << 30)
    # a good part of the lattice-like and boundary-case operands is accepted (rho within [0.5, 2],
    # |u| below the clamp); special values and out-of-range operands are not
    assert 0.15 * total["checked"] < total["accepted"] < total["checked"], total

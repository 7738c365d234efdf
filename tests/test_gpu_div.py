"""The fused kernels compute ux = jx/rho, uy = jy/rho with a shared reciprocal (div_pair in
alb_lbm.cuh) instead of two generic divisions.  Inside its accepted operand range it must equal
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


# ---- division by tau: verified shortcut or IEEE division, never an unverified shortcut ------------

from conftest import assert_bitwise  # noqa: E402
import numpy as np  # noqa: E402


def _oracle_run(nx, ny, shape, alpha, u0, tau, nsteps):
    from oracle import geometry as ogeo
    from oracle import lbm as olbm
    o = olbm.OracleTunnel(nx, ny, u0=u0, tau=tau)
    o.apply_geometry(ogeo.SHAPES[shape](), alpha)
    o.step(nsteps)
    return o


@pytest.mark.parametrize("tau", [0.58, 0.51, 0.5004, 0.75, 1.0, 1.37, 2.0])
def test_div_mode_is_verified_per_tau(built_lib, tau):
    """alb_set_params runs the exhaustive device check for a new tau; whichever mode it selects, and
    with IEEE division forced, the step is bit-identical to the oracle on all three launch paths."""
    import aerolab_lbm as al
    cases = [(320, 160, -1), (700, 130, 0), (700, 130, 1)]       # persistent small-lattice kernel, single steps, double steps
    for nx, ny, dbl in cases:
        o = _oracle_run(nx, ny, "naca4412", 9.0, 0.07, tau, 13)
        for force_ieee in (False, True):
            with al.WindTunnel(nx, ny, 0, u0=0.07, tau=tau) as t:
                t.set_double_steps(dbl)
                if force_ieee:
                    t.set_div_mode(1)
                    assert t.div_mode() == 1
                else:
                    assert t.div_mode() in (0, 1)
                t.load_shape("naca4412", alpha=9.0)
                t.step(13)
                assert_bitwise(t.populations(), o.F, f"tau={tau} {nx}x{ny} doubles={dbl} ieee={force_ieee}")


def test_reference_tau_uses_the_shortcut(built_lib):
    import aerolab_lbm as al
    with al.WindTunnel(320, 160, 0) as t:
        assert t.div_mode() == 0            # tau = 0.58 (HTML:78) passes the exhaustive check
        t.set_tau(0.6)
        assert t.div_mode() in (0, 1)
        t.set_div_mode(1)
        t.set_tau(0.58)
        assert t.div_mode() == 1            # forced stays forced
        t.set_div_mode(-1)
        assert t.div_mode() == 0


def test_random_taus_against_oracle(built_lib):
    """Random relaxation times (the reference only ever uses 0.58): whatever division mode the
    device check selects, populations equal the strict-fp32 oracle bit for bit."""
    import aerolab_lbm as al
    rng = np.random.default_rng(2024)
    modes = []
    for tau in np.concatenate([rng.uniform(0.5001, 2.0, 10), [0.5 + 2.0 ** -12, 1.9999999, 0.99999994]]):
        tau = float(np.float32(tau))
        o = _oracle_run(384, 96, "naca0012", 6.0, 0.05, tau, 7)
        with al.WindTunnel(384, 96, 0, u0=0.05, tau=tau) as t:
            t.load_shape("naca0012", alpha=6.0)
            t.step(7)
            modes.append(t.div_mode())
            assert_bitwise(t.populations(), o.F, f"tau={tau!r} mode {modes[-1]}")
    assert 0 in modes

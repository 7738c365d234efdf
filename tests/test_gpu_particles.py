"""GPU: tracer particles against the oracle's restatement with the same
counter-based random streams."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import geometry as ogeo  # noqa: E402
from oracle import lbm as olbm  # noqa: E402
from oracle.particles import Particles  # noqa: E402


@pytest.fixture(scope="module")
def al(built_lib):
    import aerolab_lbm
    return aerolab_lbm


def test_particles_match_oracle(al):
    nx, ny, n, seed = 320, 160, 600, 12345
    t = al.WindTunnel(nx, ny, 0)
    t.load_shape("naca4412", alpha=12.0)
    o = olbm.OracleTunnel(nx, ny)
    o.apply_geometry(ogeo.SHAPES["naca4412"](), 12.0)
    t.init_particles(n, seed)
    P = Particles(n, seed)
    got = t.particles()
    assert got.shape == (n, 8)
    assert np.array_equal(got[:, :4], P.table())            # initParts: identical streams, exact
    assert (got[:, 0] >= ogeo.DX0).all() and (got[:, 0] <= ogeo.DX1).all()
    respawns = 0
    for frame in range(40):
        t.step(4); o.step(4)
        dt = 16.0 if frame == 0 else 33.0
        t.step_particles(dt)
        want = P.step(dt, o.mask, o.ux, o.uy, o.u0)
        got = t.particles()
        assert np.array_equal(got[:, 7], want[:, 7]), frame  # same respawn decisions
        np.testing.assert_allclose(got[:, :7], want[:, :7], rtol=1e-12, atol=1e-15)
        respawns += int(want[:, 7].sum())
    assert respawns > 0                                      # bodies and exits were actually hit
    # moving particles drift downstream on average
    moving = got[:, 7] == 0
    assert (got[moving, 0] - got[moving, 4]).mean() > 0
    # no particle sits inside the body: its four surrounding cells cannot all be solid
    for x, y in got[:, :2]:
        fx = (x - ogeo.DX0) / (ogeo.DX1 - ogeo.DX0) * nx - 0.5
        fy = (y - ogeo.DY0) / (ogeo.DY1 - ogeo.DY0) * ny - 0.5
        ix, iy = int(max(0, min(np.floor(fx), nx - 2))), int(max(0, min(np.floor(fy), ny - 2)))
        assert not o.mask[iy:iy + 2, ix:ix + 2].all()


def test_particle_slider_and_determinism(al):
    t = al.WindTunnel(128, 64, 0)
    t.load_shape("naca0012", alpha=0.0)
    t.step(20)
    t.init_particles(100, 7)
    P = Particles(100, 7)
    t.resize_particles(160)
    P.resize(160)
    assert np.array_equal(t.particles()[:, :4], P.table())
    t.resize_particles(40)
    P.resize(40)
    assert np.array_equal(t.particles()[:, :4], P.table())
    t.step_particles(16.0)
    a = t.particles()
    t2 = al.WindTunnel(128, 64, 0)
    t2.load_shape("naca0012", alpha=0.0)
    t2.step(20)
    t2.init_particles(100, 7).resize_particles(160).resize_particles(40).step_particles(16.0)
    assert np.array_equal(a, t2.particles())               # same seed, same run
    t2.init_particles(40, 8)
    assert not np.array_equal(t2.particles()[:, :4], P.table())
    with pytest.raises(al.AerolabLbmError):
        t.init_particles(-1, 0)

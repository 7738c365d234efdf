"""GPU parity of band_lattice_kernel, the register-resident path for small lattices (the reference's
default 320x160, HTML:76): populations, macroscopic fields, momentum-exchange history across
launch chunks, clamp hits, statistics and the frame loop are bit-identical to the oracle, for
lattices that do and do not divide evenly over the SMs, with solids on every border."""
import numpy as np
import pytest

from conftest import assert_bitwise

pytestmark = pytest.mark.gpu

from oracle import geometry as ogeo  # noqa: E402
from oracle import lbm as olbm  # noqa: E402


@pytest.fixture(scope="module")
def al(built_lib):
    import aerolab_lbm
    return aerolab_lbm


@pytest.mark.parametrize("nx,ny,batches", [
    (320, 160, (2, 3, 40, 7)),          # the page's lattice: 148 CTAs of 346 cells
    (320, 161, (5, 6)),                 # the last CTA is short and holds only border cells
    (97, 31, (4, 9)),                   # 32 cells per CTA: fewer CTAs than SMs
    (700, 200, (3, 8)),                 # 946 cells per SM: does not qualify -> grid-barrier kernel
    (250, 444, (6, 5)),                 # 750 cells per SM: grid-barrier kernel again
    (200, 280, (6, 5)),                 # 379 cells per CTA
    (64, 9, (7, 2, 2)),                 # tiny: most threads are border cells that poll without using the values
])
def test_band_kernel_bitwise(al, nx, ny, batches):
    rng = np.random.default_rng(nx * ny)
    u0, tau = 0.08, 0.57
    m = (rng.random((ny, nx)) < 0.01).astype(np.uint8) * 255
    m[ny // 3:ny // 3 + 5, nx // 4:nx // 4 + 30] = 255
    m[0, 5:9] = 255; m[ny - 1, 10:12] = 255; m[ny // 2, 0] = 255; m[ny // 2 + 1, nx - 1] = 255
    t = al.WindTunnel(nx, ny, 0, u0=u0, tau=tau)
    t.set_mask(m)
    o = olbm.OracleTunnel(nx, ny, u0=u0, tau=tau)
    o.set_mask(m)
    total = 0
    for n in batches:
        t.step(n); o.step(n)
        total += n
        assert_bitwise(t.populations(), o.F, f"{nx}x{ny} after {total} steps")
        for a, b, name in zip(t.macro(), (o.rho, o.ux, o.uy), ("rho", "ux", "uy")):
            assert_bitwise(a, b, name)
    assert np.array_equal(t.me_history(total), np.array(o.me_hist, dtype=np.int64))
    assert t.clamp_hits() == o.clamp_hits
    st = t.update_stats(); o.update_fields()
    assert st["cpMin"] == o.cp_min and st["cpMax"] == o.cp_max
    # a single step afterwards (streaming kernels) continues the momentum-exchange bookkeeping
    t.step(1); o.step(1)
    assert_bitwise(t.populations(), o.F, "single step after band batches")
    assert np.array_equal(t.me_history(total + 1), np.array(o.me_hist, dtype=np.int64))
    t.close()


def test_band_kernel_long_batch_crosses_the_history_ring(al):
    """One alb_step call of more steps than half the momentum-exchange ring: several launches."""
    nx, ny, n = 320, 160, 4500
    t = al.WindTunnel(nx, ny, 0)
    t.load_shape("naca0012", alpha=5.0)
    o = olbm.OracleTunnel(nx, ny)
    o.apply_geometry(ogeo.SHAPES["naca0012"](), 5.0)
    t.step(n); o.step(n)
    assert_bitwise(t.populations(), o.F, f"after {n} steps in one call")
    hist = t.me_history(4095)
    assert np.array_equal(hist, np.array(o.me_hist[-4095:], dtype=np.int64))
    t.close()


def test_band_and_grid_barrier_kernels_agree(al, monkeypatch):
    nx, ny = 320, 160
    monkeypatch.setenv("AEROLAB_LBM_BAND", "0")
    a = al.WindTunnel(nx, ny, 0)
    monkeypatch.delenv("AEROLAB_LBM_BAND")
    b = al.WindTunnel(nx, ny, 0)
    for t in (a, b):
        t.load_shape("naca4412", alpha=14.0)
    fa = a.run_frames(30); fb = b.run_frames(30)
    for k in fa:
        x, y = fa[k], fb[k]
        assert np.array_equal(np.isnan(x), np.isnan(y)) and np.array_equal(x[~np.isnan(x)], y[~np.isnan(y)]), k
    assert_bitwise(a.populations(), b.populations(), "band vs grid-barrier kernel after 30 frames")
    assert b.launch_count() != a.launch_count()        # they really took different paths
    a.close(); b.close()

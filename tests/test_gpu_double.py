"""GPU parity of the two-steps-per-pass path (march2_kernel + the list-driven two-pass path).

`set_double_steps(1)` forces it on lattices far below the automatic threshold, so the oracle
finishes in seconds.  The bar is the same as for single steps: populations, macroscopic fields,
momentum-exchange integers, clamp-hit counts and the statistics of the final state are
bit-identical to the strict-fp32 oracle.
"""
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


def compare_state(t, o, what):
    assert_bitwise(t.populations(), o.F, what + " populations")
    rho, ux, uy = t.macro()
    assert_bitwise(rho, o.rho, what + " rho")
    assert_bitwise(ux, o.ux, what + " ux")
    assert_bitwise(uy, o.uy, what + " uy")


def compare_diagnostics(t, o, total, what):
    assert np.array_equal(t.me_history(total), np.array(o.me_hist, dtype=np.int64)), what
    st = t.update_stats(); o.update_fields()
    assert st["cpMin"] == o.cp_min and st["cpMax"] == o.cp_max, what
    assert st["maxS"] == pytest.approx(o.max_s, rel=1e-14), what
    f = t.forces(); w = o.compute_forces()
    if w is None:
        assert not f["any"], what
    else:
        assert (f["surf"], f["rev"]) == (w["surf"], w["rev"]), what
        assert f["CL_raw"] == pytest.approx(w["CL_raw"], rel=1e-11, abs=1e-11), what
    assert t.clamp_hits() == o.clamp_hits, what


@pytest.mark.parametrize("nx,ny,shape,alpha,batches", [
    (320, 160, "naca0012", 5.0, (3, 4, 21)),          # one column segment (384 columns: a single task can be deep)
    (1200, 300, "naca4412", 10.0, (5, 12, 9)),        # eight column segments, graph replay
    (333, 171, "naca2412", 12.0, (3, 7)),             # padded pitch
    (2048, 520, "clark_y", 6.0, (11, 10)),            # fifteen column segments
    (1536, 66, "naca0012", 0.0, (4, 5)),              # fewer rows than one segment
    (130, 7, "naca0012", 3.0, (3, 3, 3)),             # hardly any deep task
    (700, 5, "naca0012", 3.0, (9,)),                  # three deep rows at most
])
def test_double_steps_bitwise(al, nx, ny, shape, alpha, batches):
    t = al.WindTunnel(nx, ny, 0)
    t.set_double_steps(1)
    t.load_shape(shape, alpha=alpha)
    o = olbm.OracleTunnel(nx, ny)
    o.apply_geometry(ogeo.SHAPES[shape](), alpha)
    total = 0
    for n in batches:
        t.step(n); o.step(n)
        total += n
        compare_state(t, o, f"{nx}x{ny} {shape} after {total} steps")
    compare_diagnostics(t, o, total, f"{nx}x{ny} {shape}")
    t.close()


def test_double_steps_random_masks(al):
    rng = np.random.default_rng(77)
    sizes = [(int(rng.integers(260, 1400)), int(rng.integers(6, 300))) for _ in range(8)]
    for k, (nx, ny) in enumerate(sizes):
        u0 = float(rng.uniform(0.03, 0.1))
        tau = float(rng.uniform(0.51, 1.2))
        # sparse obstacles leave deep regions between them; dense ones leave none
        dens = float(rng.choice([0.0, 0.0002, 0.002, 0.05]))
        m = (rng.random((ny, nx)) < dens).astype(np.uint8) * 255
        t = al.WindTunnel(nx, ny, 0, u0=u0, tau=tau)
        t.set_double_steps(1)
        o = olbm.OracleTunnel(nx, ny, u0, tau)
        t.set_mask(m); o.set_mask(m)
        total = 0
        for n in (int(rng.integers(3, 8)), int(rng.integers(9, 30)), 1, 2):
            t.step(n); o.step(n)
            total += n
        what = f"case {k}: {nx}x{ny} u0={u0:.3f} tau={tau:.3f} dens={dens}"
        compare_state(t, o, what)
        compare_diagnostics(t, o, total, what)
        t.close()


def test_double_steps_clamps_counted_once(al):
    """Clamp hits in deep cells are counted by the fused kernel, in shallow cells by the two-pass
    path, and never twice (pass 1 also recomputes deep neighbours of shallow tasks)."""
    nx, ny = 900, 160
    m = np.zeros((ny, nx), np.uint8)
    m[20:140, 300:304] = 255
    t = al.WindTunnel(nx, ny, 0, u0=0.25, tau=0.505)
    t.set_double_steps(1)
    t.set_mask(m)
    o = olbm.OracleTunnel(nx, ny, 0.25, 0.505)
    o.set_mask(m)
    for n in (151, 150, 100):
        t.step(n); o.step(n)
        assert t.clamp_hits() == o.clamp_hits
    assert o.clamp_hits > 0
    compare_state(t, o, "clamped, double steps")
    t.close()


def test_double_and_single_modes_interleave(al):
    """Switching the mode between batches, parameter changes and mask changes in between."""
    nx, ny = 1100, 200
    t = al.WindTunnel(nx, ny, 0)
    t.load_shape("naca2412", alpha=6.0)
    o = olbm.OracleTunnel(nx, ny)
    o.apply_geometry(ogeo.SHAPES["naca2412"](), 6.0)
    total = 0
    for mode, n in ((1, 7), (0, 4), (1, 3), (1, 10), (0, 9), (1, 12)):
        t.set_double_steps(mode)
        t.step(n); o.step(n)
        total += n
    compare_state(t, o, "interleaved modes")
    t.set_double_steps(1)
    t.set_params(0.08, 0.62); o.u0, o.tau = 0.08, 0.62
    t.step(6); o.step(6)
    t.set_alpha(11.0); o.apply_geometry(ogeo.SHAPES["naca2412"](), 11.0)
    t.step(13); o.step(13)
    total += 19
    compare_state(t, o, "after parameter and geometry changes")
    compare_diagnostics(t, o, total, "interleaved")
    t.close()


def test_double_steps_frame_loop(al):
    """The page's frame loop (4 steps + statistics per frame) with double steps inside."""
    nx, ny = 1100, 200
    a = al.WindTunnel(nx, ny, 0)
    b = al.WindTunnel(nx, ny, 0)
    b.set_double_steps(1)
    for t in (a, b):
        t.load_shape("naca4412", alpha=9.0)
    ra = a.run_frames(6, None, 4, 3)
    rb = b.run_frames(6, None, 4, 3)
    for k in ra:
        assert np.array_equal(ra[k], rb[k], equal_nan=True), k
    assert_bitwise(a.populations(), b.populations(), "frame loop populations")
    a.close(); b.close()


@pytest.mark.parametrize("splits", [[(0, 100), (100, 100)], [(0, 60), (60, 90), (150, 50)]])
def test_double_steps_local_slabs(al, splits):
    """Slabs with the halo pushed by the two-pass path (edge rows are always shallow)."""
    nx, ny = 1100, 200
    whole = al.WindTunnel(nx, ny, 0)
    whole.load_shape("naca4412", alpha=10.0)
    slabs = [al.WindTunnel(nx, ny, 0, y0=y0, ny_local=n) for y0, n in splits]
    for s in slabs:
        s.set_double_steps(1)
        s.load_shape("naca4412", alpha=10.0)
    for k, s in enumerate(slabs):
        s.connect_local(slabs[k - 1] if k > 0 else None, slabs[k + 1] if k + 1 < len(slabs) else None)
    nsteps = 0
    for n in (3, 4, 7):
        for s in slabs:
            s.step(n)
        nsteps += n
    for s in slabs:
        s.sync()
    whole.step(nsteps)
    assert_bitwise(np.concatenate([s.populations() for s in slabs], 1), whole.populations(), "slab populations")
    wm = whole.macro()
    for k in range(3):
        assert_bitwise(np.concatenate([s.macro()[k] for s in slabs], 0), wm[k], f"slab macro {k}")
    me = sum(s.me_history(1)[0] for s in slabs)
    assert np.array_equal(me, whole.me_history(1)[0])


def test_double_steps_restart(al):
    """dump_state / load_state in the middle of a double-step run (the all-solid tasks of the other
    ping-pong buffer are stale after set_populations and must be copied again)."""
    nx, ny = 1300, 220
    t = al.WindTunnel(nx, ny, 0)
    t.set_double_steps(1)
    t.load_shape("naca4412", alpha=12.0)
    o = olbm.OracleTunnel(nx, ny)
    o.apply_geometry(ogeo.SHAPES["naca4412"](), 12.0)
    t.step(10); o.step(10)
    st = t.dump_state()
    t.step(6)                      # diverge ...
    t.load_state(st)               # ... and come back
    t.step(8); o.step(8)
    compare_state(t, o, "restart in double mode")
    t2 = al.WindTunnel(nx, ny, 0)
    t2.set_double_steps(1)
    t2.load_shape("naca4412", alpha=12.0)
    t2.load_state(st)
    t2.step(8)
    assert_bitwise(t2.populations(), o.F, "fresh tunnel restarted from the dump")
    t.close(); t2.close()


@pytest.mark.parametrize("nx,ny,dens", [
    (256, 40, 0.0),                     # two tasks per row: inlet segment, outlet segment and a clipped one between
    (384, 70, 0.001),
    (512, 33, 0.0),
    (1280, 130, 0.0005),
    (2048, 48, 0.0),
    (260, 40, 0.0),                     # padded rows: the last task holds three plain cells, the outlet cell and padding
    (500, 45, 0.001),                   # the last segment straddles the last two tasks
    (2000, 50, 0.0),
])
def test_double_steps_inlet_and_outlet_columns(al, nx, ny, dens):
    """Widths that are a multiple of 4: the inlet column x = 0 and the outlet column x = nx-1 are part
    of the fused kernel's domain (alb_march.cu patches the one special cell).
    A perturbed state makes the outlet copy (HTML:301-312) visible; solids next to and on the two
    columns leave some edge tasks to the list-driven passes."""
    rng = np.random.default_rng(nx * 131 + ny)
    u0, tau = 0.09, 0.58
    m = (rng.random((ny, nx)) < dens).astype(np.uint8) * 255
    if ny >= 40:
        m[ny // 4, 1] = 255                  # touches the inlet column
        m[ny // 4 + 6, 0] = 255              # on it
        m[ny // 2, nx - 2] = 255             # the cell the outlet copies from
        m[ny // 2 + 7, nx - 1] = 255         # on the outlet column
        m[3 * ny // 4:3 * ny // 4 + 3, nx // 2:nx // 2 + 9] = 255
    F, _, _, _ = olbm.init(nx, ny, u0)
    F *= (np.float32(1.0) + np.float32(4e-3) * (rng.random(F.shape, dtype=np.float32) - np.float32(0.5)))
    t = al.WindTunnel(nx, ny, 0, u0=u0, tau=tau)
    t.set_double_steps(1)
    plan = t.step2_plan()
    assert plan["nseg"] == 2 + max(0, -(-(nx - 248) // 120)), plan
    t.set_mask(m)
    t.set_populations(F)
    G = np.empty_like(F)
    rho = np.empty((ny, nx), np.float32); ux = np.empty_like(rho); uy = np.empty_like(rho)
    me, hits, total = [], 0, 0
    for n in (2, 7, 12, 1, 4):
        t.step(n)
        for _ in range(n):
            fx, fy, h = olbm.step(m, F, G, rho, ux, uy, tau, u0)
            F, G = G, F
            me.append((fx, fy)); hits += h
        total += n
        assert_bitwise(t.populations(), F, f"{nx}x{ny} populations after {total} steps")
    r, x, y = t.macro()
    assert_bitwise(r, rho, "rho"); assert_bitwise(x, ux, "ux"); assert_bitwise(y, uy, "uy")
    assert np.array_equal(t.me_history(total), np.array(me, dtype=np.int64))
    assert t.clamp_hits() == hits
    # the statistics fused into the double step that ends a frame see the two columns too
    t.close()
    pair = []
    for mode in (0, 1):
        s = al.WindTunnel(nx, ny, 0, u0=u0, tau=tau)
        s.set_double_steps(mode)
        s.set_mask(m)
        s.set_populations(F)
        pair.append((s, s.run_frames(5, None, 4, 2)))
    (a, ra), (b, rb) = pair
    for k in ra:
        assert np.array_equal(ra[k], rb[k], equal_nan=True), k
    assert_bitwise(a.populations(), b.populations(), "frame loop populations")
    a.close(); b.close()


def test_double_steps_outlet_column_extreme_cell(al):
    """The statistics' arg-max lies ON the outlet column / the rho extrema on it: the fused kernel must
    account for the patched cells exactly like the macroscopic pass does."""
    nx, ny, u0, tau = 512, 64, 0.06, 0.6
    F, _, _, _ = olbm.init(nx, ny, u0)
    # a fast, dense blob left of the outlet: two steps later its copy sits in the outlet column
    F[:, 20:24, nx - 4:nx - 1] *= np.float32(1.05)
    F[1, 20:24, nx - 4:nx - 1] *= np.float32(1.6)
    F[:, 40:42, 1:3] *= np.float32(0.97)
    a = al.WindTunnel(nx, ny, 0, u0=u0, tau=tau)
    b = al.WindTunnel(nx, ny, 0, u0=u0, tau=tau)
    a.set_double_steps(0); b.set_double_steps(1)
    for t in (a, b):
        t.set_populations(F)
    ra = a.run_frames(4, None, 2, 1)
    rb = b.run_frames(4, None, 2, 1)
    for k in ra:
        assert np.array_equal(ra[k], rb[k], equal_nan=True), (k, ra[k], rb[k])
    assert_bitwise(a.populations(), b.populations(), "populations")
    a.close(); b.close()

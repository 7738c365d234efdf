"""GPU: the control surface on a slab-decomposed lattice (in-process slabs on one device; the same
scenarios run across processes / devices in tests/_dist_gpu_worker.py).

Round-1 advisor findings covered here: the lazy macroscopic pass must not see ghost rows that a
neighbour has already overwritten (mid-run U0 / alpha / tau changes, macro() between batches), a
reset must not race with a neighbour's late halo push, 1-row border slabs must push their
equilibrium row, and closing must not free memory a neighbour still writes to.
"""
import numpy as np
import pytest

from conftest import assert_bitwise

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def al(built_lib):
    import aerolab_lbm
    return aerolab_lbm


def same_series(a, b):
    for k in a:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.shape == y.shape, k
        assert np.array_equal(np.isnan(x), np.isnan(y)), k
        assert np.array_equal(x[~np.isnan(x)], y[~np.isnan(y)]), (k, x, y)


@pytest.mark.parametrize("nx,ny,double,devs", [(640, 301, 0, [0, 0, 0]), (1400, 260, 1, [0, 0])])
def test_run_frames_on_slabs_equals_whole_lattice(al, nx, ny, double, devs):
    """alb_frames_enqueue on every slab + combine_frame_partials == alb_run_frames on one GPU, with
    slider changes between frames (HTML:956-959) and the 3-frame force cadence (HTML:914)."""
    whole = al.WindTunnel(nx, ny, 0)
    whole.set_double_steps(double)
    whole.load_shape("naca4412", alpha=11.0)
    multi = al.LocalMultiTunnel(nx, ny, devs)
    for t in multi.slabs:
        t.set_double_steps(double)
    multi.load_shape("naca4412", alpha=11.0)
    nframes = 14
    controls = np.tile(np.array([0.06, 0.58]), (nframes, 1))
    controls[5:, 0] = 0.07
    controls[9:, 1] = 0.6
    a = whole.run_frames(nframes, controls=controls)
    b = multi.run_frames(nframes, controls=controls)
    same_series(a, b)
    # a second batch continues the EMAs and the frame counter
    same_series(whole.run_frames(4, controls=controls[-4:]), multi.run_frames(4, controls=controls[-4:]))
    multi.sync()
    assert_bitwise(multi.populations(), whole.populations(), "populations after the frame loops")
    multi.close()
    whole.close()


@pytest.mark.parametrize("double", [0, 1])
def test_mid_run_changes_on_slabs(al, double):
    """U0, tau and alpha change between batches, macro() is read between batches: every one of them
    runs the lazy macroscopic pass over the PREVIOUS state, whose ghost rows the neighbours must
    not have overwritten yet."""
    nx, ny = (1300, 240) if double else (512, 200)
    whole = al.WindTunnel(nx, ny, 0)
    whole.set_double_steps(double)
    whole.load_shape("naca2412", alpha=6.0)
    multi = al.LocalMultiTunnel(nx, ny, [0, 0, 0])
    for t in multi.slabs:
        t.set_double_steps(double)
    multi.load_shape("naca2412", alpha=6.0)
    script = [("step", 9), ("u0", 0.08), ("step", 6), ("macro", None), ("step", 5), ("tau", 0.7), ("step", 8),
              ("alpha", 12.0), ("step", 7), ("macro", None), ("step", 4)]
    for op, v in script:
        if op == "step":
            whole.step(v); multi.step(v)
        elif op == "u0":
            whole.set_u0(v); multi.set_params(v, whole.params()[1])
        elif op == "tau":
            whole.set_tau(v); multi.set_params(whole.params()[0], v)
        elif op == "alpha":
            whole.set_alpha(v); multi.set_alpha(v)
        else:
            for x, y in zip(multi.macro(), whole.macro()):
                assert_bitwise(x, y, "macro between batches")
    multi.sync()
    assert_bitwise(multi.populations(), whole.populations(), "populations after mid-run changes")
    for x, y in zip(multi.macro(), whole.macro()):
        assert_bitwise(x, y, "final macro")
    multi.close()
    whole.close()


def test_reset_on_slabs(al):
    nx, ny = 512, 160
    whole = al.WindTunnel(nx, ny, 0)
    whole.load_shape("naca0012", alpha=9.0)
    multi = al.LocalMultiTunnel(nx, ny, [0, 0])
    multi.load_shape("naca0012", alpha=9.0)
    for n, u0 in ((12, 0.05), (7, 0.09), (10, 0.06)):      # even and odd step counts before a reset
        whole.step(n); multi.step(n)
        whole.reset(u0); multi.reset(u0)
        whole.step(5); multi.step(5)
        multi.sync()
        assert_bitwise(multi.populations(), whole.populations(), f"after reset to U0={u0}")
    multi.close()
    whole.close()


def test_one_row_border_slabs_push_their_equilibrium_row(al):
    """A bottom / top slab of a single row is all-equilibrium; after a U0 change WITHOUT a reset its
    neighbours must see the new border values."""
    nx, ny = 384, 64
    whole = al.WindTunnel(nx, ny, 0)
    whole.load_shape("naca0012", alpha=4.0)
    splits = [(0, 1), (1, 62), (63, 1)]
    slabs = [al.WindTunnel(nx, ny, 0, y0=y0, ny_local=n) for y0, n in splits]
    for s in slabs:
        s.load_shape("naca0012", alpha=4.0)
    for k, s in enumerate(slabs):
        s.connect_local(slabs[k - 1] if k > 0 else None, slabs[k + 1] if k + 1 < len(slabs) else None)
    for u0, n in ((0.06, 6), (0.09, 9), (0.04, 5)):
        whole.set_u0(u0)
        for s in slabs:
            s.set_u0(u0)
        whole.step(n)
        for _ in range(n):
            for s in slabs:
                s.step(1)
    for s in slabs:
        s.sync()
    assert_bitwise(np.concatenate([s.populations() for s in slabs], 1), whole.populations(), "1-row border slabs")
    for s in slabs:
        s.sync()
    for s in slabs:
        s.close()
    whole.close()


def test_state_hash_adds_up_over_slabs(al):
    nx, ny = 700, 150
    whole = al.WindTunnel(nx, ny, 0)
    whole.load_shape("naca4412", alpha=8.0)
    multi = al.LocalMultiTunnel(nx, ny, [0, 0, 0])
    multi.load_shape("naca4412", alpha=8.0)
    whole.step(33); multi.step(33); multi.sync()
    hw = whole.state_hash()
    assert np.array_equal(hw, al.state_hash_numpy(whole.populations(), nx))
    hs = np.zeros(9, np.uint64)
    with np.errstate(over="ignore"):
        for t in multi.slabs:
            hs += t.state_hash()
    assert np.array_equal(hs, hw)
    # position dependence: swapping two rows changes the checksum, a plain sum would not notice
    F = whole.populations()
    F[:, [40, 41]] = F[:, [41, 40]]
    assert not np.array_equal(al.state_hash_numpy(F, nx), hw)
    # band getter
    assert_bitwise(whole.population_rows(40, 5), whole.populations()[:, 40:45], "population_rows")
    multi.close()
    whole.close()


@pytest.mark.parametrize("devs", [[0, 0], [0, 0, 0]])
def test_field_modes_on_slabs_equal_whole_lattice(al, devs):
    """renderField (HTML:530-545) per slab: speed and Cp are cell-local, the vorticity taps of a
    slab's first / last row (HTML:411-418) read the neighbour's edge row through the ghost rows."""
    nx, ny = 640, 301
    whole = al.WindTunnel(nx, ny, 0)
    whole.load_shape("naca4412", alpha=12.0)
    multi = al.LocalMultiTunnel(nx, ny, devs)
    multi.load_shape("naca4412", alpha=12.0)
    whole.step(150); multi.step(150)
    sw, sm = whole.update_stats(), multi.update_stats()
    assert sw["cpMin"] == sm["cpMin"] and sw["cpMax"] == sm["cpMax"] and sw["maxS"] == pytest.approx(sm["maxS"], rel=1e-14)
    multi._sticky["maxS"] = sw["maxS"]; multi._push_stats()      # maxS: device hypot vs host combine, 1e-14
    for mode in ("speed", "cp", "vort"):
        assert_bitwise(multi.field(mode), whole.field(mode), f"field {mode}")
        assert np.array_equal(multi.rgba(mode), whole.rgba(mode)), mode
    # vorticity without the neighbours' rows is refused, not silently clamped
    multi.step(1)
    with pytest.raises(al.AerolabLbmError):
        multi.slabs[0].field("vort")
    multi.slabs[0].field("speed")
    multi.close()
    whole.close()

"""CPU tests of the oracle itself: pins against the committed golden fixtures, the surveyor's
probe values and the second, independent NumPy restatement."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import assert_bitwise
from oracle import geometry as ogeo
from oracle import lbm as olbm
from oracle import lbm_numpy as onp

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("case", GOLD["masks"], ids=lambda c: f"{c['shape']}-{c['alpha']}-{c['nx']}")
def test_mask_hashes(case):
    _, _, m = ogeo.build_geometry(ogeo.SHAPES[case["shape"]](), case["alpha"], case["nx"], case["ny"])
    assert int((m > 0).sum()) == case["solid"]
    assert sha(m) == case["sha256"]


def test_surveyor_probe_pins():
    """Independent throw-away restatement by the surveyor."""
    _, _, m = ogeo.build_geometry(ogeo.SHAPES["naca0012"](), 5.0, 320, 160)
    assert int((m > 0).sum()) == 2470
    assert sha(m).startswith("d0002826024cfacf")
    ys, xs = np.nonzero(m)
    assert (xs.min(), xs.max(), ys.min(), ys.max()) == (74, 244, 67, 90)
    _, _, m = ogeo.build_geometry(ogeo.SHAPES["naca2412"](), 6.0, 320, 160)
    assert int((m > 0).sum()) == 2463
    for alpha, n in ((0.0, 101042), (20.0, 101007)):
        _, _, m = ogeo.build_geometry(ogeo.SHAPES["naca0012"](), alpha, 2048, 1024)
        assert int((m > 0).sum()) == n


def test_shapes_structure():
    pts = ogeo.naca4(0, 0, 12, 50)
    assert len(pts) == 101
    assert pts[0][0] == pytest.approx(1.0) and pts[-1][0] == pytest.approx(1.0)
    assert abs(pts[0][1]) < 1e-15 and abs(pts[-1][1]) < 1e-15      # closed TE (-0.1036 coefficient)
    assert pts[50] == [0.0, 0.0]
    cy = ogeo.clark_y()
    assert len(cy) == 35 and cy[0] == [1.0, 0.0044] and cy[-1] == [1.0, -0.0044]   # open TE
    xp, yp = ogeo.panelise(ogeo.rotate(pts, 5.0))
    assert len(xp) == len(yp) == ogeo.NP + 1


def test_rotation_pivot_and_sign():
    # positive alpha pitches the nose up: the TE (x=1) moves DOWN, rotation about (0.25, 0)
    (x, y), = ogeo.rotate([[1.0, 0.0]], 10.0)
    assert y < 0 and x < 1.0
    (x, y), = ogeo.rotate([[0.25, 0.0]], 33.0)
    assert (x, y) == (0.25, 0.0)


def test_open_te_slit_rows():
    pts = ogeo.round_coords(GOLD["parser"]["naca0012_selig_test_main"]["coords"])
    _, _, m = ogeo.build_geometry(pts, 0.0, 2048, 1024)
    assert m[511].sum() == 0 and m[512].sum() == 0 and m[510].sum() > 0 and m[513].sum() > 0


@pytest.mark.parametrize("nx,ny,shape,alpha,u0,tau", [(320, 160, "naca0012", 5.0, 0.06, 0.58),
                                                      (97, 45, "clark_y", 12.0, 0.1, 0.53)])
def test_c_and_numpy_restatements_agree_bitwise(nx, ny, shape, alpha, u0, tau):
    o = olbm.OracleTunnel(nx, ny, u0, tau)
    o.apply_geometry(ogeo.SHAPES[shape](), alpha)
    F = onp.init(nx, ny, u0)
    assert_bitwise(F, o.F, "init")
    for s in range(40):
        o.step(1)
        F, rho, ux, uy = onp.step(o.mask, F, tau, u0)
        assert_bitwise(F, o.F, f"F step {s}")
        assert_bitwise(rho, o.rho, f"rho step {s}")
        assert_bitwise(ux, o.ux, f"ux step {s}")
        assert_bitwise(uy, o.uy, f"uy step {s}")


def test_random_mask_with_borders_c_vs_numpy():
    rng = np.random.default_rng(11)
    nx, ny = 64, 40
    m = (rng.random((ny, nx)) < 0.1).astype(np.uint8) * 255
    m[0, 3:9] = 255; m[-1, 20:30] = 255; m[5:9, 0] = 255; m[10:20, -1] = 255
    o = olbm.OracleTunnel(nx, ny)
    o.set_mask(m)
    F = onp.init(nx, ny, 0.06)
    for s in range(25):
        o.step(1)
        F, rho, ux, uy = onp.step(m, F, 0.58, 0.06)
        assert_bitwise(F, o.F, f"F step {s}")
        assert_bitwise(ux, o.ux, f"ux step {s}")


def test_config0_golden_field_dump():
    """configs[0]: NACA 0012 alpha=5, 320x160, 1,000 steps."""
    g = GOLD["config0"]
    o = olbm.OracleTunnel(320, 160)
    o.apply_geometry(ogeo.SHAPES["naca0012"](), 5.0)
    for s in range(1, 1001):
        o.step(1)
        if str(s) in g["hashes"]:
            h = g["hashes"][str(s)]
            assert (sha(o.F), sha(o.rho), sha(o.ux), sha(o.uy)) == (h["f"], h["rho"], h["ux"], h["uy"]), s
    d = np.load(os.path.join(HERE, "golden", "config0_step1000.npz"))
    assert_bitwise(o.rho, d["rho"], "rho")
    assert_bitwise(o.ux, d["ux"], "ux")
    assert_bitwise(o.uy, d["uy"], "uy")
    f = o.compute_forces()
    assert f["CL_raw"] == pytest.approx(g["CL_raw"], rel=1e-13) and f["surf"] == g["surf"] == 390
    assert [list(map(int, v)) for v in o.me_hist[:8]] == g["me_first8"]
    assert o.clamp_hits == 0
    # surveyor probe values
    assert f["CL_raw"] == pytest.approx(0.5854, abs=5e-5) and f["CD_raw"] == pytest.approx(0.1364, abs=5e-5)
    assert float(o.rho.min()) == pytest.approx(0.99544, abs=1e-5)
    assert float(o.rho.max()) == pytest.approx(1.01013, abs=1e-5)
    assert olbm.total_mass(o.F) / (320 * 160) - 1 == pytest.approx(8.8e-4, abs=2e-5)


def test_oracle_slabs_equal_whole():
    """The slab view of the oracle (used by the gloo test) reproduces the whole-lattice step."""
    nx, ny = 96, 60
    o = olbm.OracleTunnel(nx, ny)
    o.apply_geometry(ogeo.SHAPES["naca4412"](), 10.0)
    cuts = [(0, 25), (25, 35)]
    slabs = []
    for y0, n in cuts:
        F, rho, ux, uy = olbm.init(nx, n + 2, 0.06)
        m = np.zeros((n + 2, nx), np.uint8)
        lo, hi = max(0, y0 - 1), min(ny, y0 + n + 1)
        m[lo - (y0 - 1):hi - (y0 - 1)] = o.mask[lo:hi]
        slabs.append(dict(y0=y0, n=n, F=F, G=F.copy(), rho=rho, ux=ux, uy=uy, m=m))
    for _ in range(30):
        o.step(1)
        for s in slabs:
            olbm.step(s["m"], s["F"], s["G"], s["rho"], s["ux"], s["uy"], 0.58, 0.06, ny_global=ny,
                      gy0=s["y0"] - 1, j0=1, j1=s["n"] + 1)
            s["F"], s["G"] = s["G"], s["F"]
        a, b = slabs
        b["F"][:, 0] = a["F"][:, a["n"]]        # ghost below b <- top row of a
        a["F"][:, a["n"] + 1] = b["F"][:, 1]    # ghost above a <- bottom row of b
    got = np.concatenate([s["F"][:, 1:s["n"] + 1] for s in slabs], axis=1)
    assert_bitwise(got, o.F, "oracle slabs")


def test_forces_and_stats_small_known_case():
    """Hand-checkable pressure force: one solid cell in uniform rho = 1 gives zero net force,
    4 faces; a rho bump on one side gives the expected sign."""
    nx, ny = 8, 6
    m = np.zeros((ny, nx), np.uint8); m[3, 4] = 255
    rho = np.ones((ny, nx), np.float32); ux = np.full((ny, nx), 0.06, np.float32)
    out = olbm.forces_raw(m, rho, ux)
    assert tuple(out) == (0.0, 0.0, 1.0, 4.0, 0.0)
    rho[3, 3] = 1.3       # higher pressure on the left face pushes the body in +x
    ux[2, 4] = -0.01      # one reversed-flow face
    fx, fy, any_, surf, rev = olbm.forces_raw(m, rho, ux)
    assert fx == pytest.approx((1.3 - 1.0) / 3, rel=1e-6) and fy == 0.0 and rev == 1

"""ctypes front end of the C oracle + a frame-loop mirror (ORACLE, tests only).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Pinned against the reference's own source executed by tests/refexec (see oracle/__init__.py).


``OracleTunnel`` restates the reference's host-side driver around the step:
``initSim`` (HTML:492-500), ``applyGeometry`` (HTML:579-586, no flow reset),
``simStep`` ping-pong (HTML:510-525), ``updateFieldsFromMacro`` (HTML:596-614),
``computeForces`` with its EMAs (HTML:650-700), Reynolds number and the stall
text (HTML:862-885).
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from . import build as _build
from . import geometry as geo

_lib = None

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i64p = C.POINTER(C.c_int64)


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        L = C.CDLL(path)
        L.orc_weights.argtypes = [_f32p]
        L.orc_init.argtypes = [C.c_int, C.c_int, C.c_double, _f32p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_step.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _f32p, _f32p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                               _i64p, _i64p, _i64p]
        L.orc_run.argtypes = [C.c_int, C.c_int, _u8p, _f32p, _f32p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_float, C.c_float, C.c_int, C.c_void_p, _i64p]
        L.orc_run.restype = C.c_int
        L.orc_field_stats.argtypes = [C.c_int, C.c_int, _u8p, _f32p, _f32p, _f32p, C.c_double,
                                      C.c_void_p, C.c_void_p, C.c_void_p, _f64p]
        L.orc_forces.argtypes = [C.c_int, C.c_int, _u8p, _f32p, _f32p, _f64p]
        L.orc_render_scalar.argtypes = [C.c_int, C.c_int, _u8p, _f32p, _f32p, _f32p, C.c_int,
                                        C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _f32p]
        L.orc_rgba.argtypes = [C.c_int, C.c_int, _u8p, _f32p, C.c_int, _u8p]
        L.orc_total_mass.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
        L.orc_total_mass.restype = C.c_double
        _lib = L
    return _lib


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def set_threads(n: int) -> None:
    """OpenMP thread count for the oracle's row loops (libgomp)."""
    try:
        gomp = C.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(int(n))
    except OSError:
        os.environ["OMP_NUM_THREADS"] = str(n)


def init(nx, nrows, u0):
    F = np.empty((9, nrows, nx), np.float32)
    rho = np.empty((nrows, nx), np.float32)
    ux = np.empty((nrows, nx), np.float32)
    uy = np.empty((nrows, nx), np.float32)
    lib().orc_init(nx, nrows, float(u0), F, _vp(rho), _vp(ux), _vp(uy))
    return F, rho, ux, uy


def step(mask, src, dst, rho, ux, uy, tau, u0, ny_global=None, gy0=0, j0=0, j1=None):
    """One step; returns (me_fx, me_fy, clamp_hits) of this step (fixed point 2^-40)."""
    _, nrows, nx = src.shape
    if ny_global is None:
        ny_global = nrows
    if j1 is None:
        j1 = nrows
    fx, fy, hits = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    lib().orc_step(nx, ny_global, gy0, nrows, j0, j1, mask, src, dst, _vp(rho), _vp(ux), _vp(uy),
                   np.float32(tau), np.float32(u0), C.byref(fx), C.byref(fy), C.byref(hits))
    return fx.value, fy.value, hits.value


def field_stats(mask, rho, ux, uy, u0, want_fields=False):
    ny, nx = mask.shape
    out = np.zeros(3, np.float64)
    U = V = Cp = None
    if want_fields:
        U = np.empty((ny, nx), np.float32)
        V = np.empty((ny, nx), np.float32)
        Cp = np.empty((ny, nx), np.float32)
    lib().orc_field_stats(nx, ny, mask, rho, ux, uy, float(u0), _vp(U), _vp(V), _vp(Cp), out)
    return out, U, V, Cp


def forces_raw(mask, rho, ux):
    ny, nx = mask.shape
    out = np.zeros(5, np.float64)
    lib().orc_forces(nx, ny, mask, rho, ux, out)
    return out


def render_scalar(mask, rho, ux, uy, mode, u0, max_s, cp_min, cp_max, vort_scale=0.06):
    ny, nx = mask.shape
    t = np.empty((ny, nx), np.float32)
    lib().orc_render_scalar(nx, ny, mask, rho, ux, uy, int(mode), np.float32(u0), np.float32(max_s),
                            np.float32(cp_min), np.float32(cp_max), np.float32(vort_scale), t)
    return t


def rgba(mask, t, mode):
    ny, nx = mask.shape
    out = np.empty((ny, nx, 4), np.uint8)
    lib().orc_rgba(nx, ny, mask, t, int(mode), out)
    return out


def total_mass(F, j0=0, j1=None):
    _, nrows, nx = F.shape
    if j1 is None:
        j1 = nrows
    return lib().orc_total_mass(nx, nrows, j0, j1, F)


ME_SCALE = float(2 ** 40)


class OracleTunnel:
    """Mirror of the reference page's simulation state and host-side driver."""

    TAU = 0.58            # HTML:78
    VORT_SCALE = 0.06     # HTML:528
    FORCE_EVERY = 12      # HTML:80, 914: every 3rd frame of 4 steps

    def __init__(self, nx=320, ny=160, u0=0.06, tau=TAU):
        self.nx, self.ny = nx, ny
        self.u0 = float(u0)          # JS double
        self.tau = float(tau)
        self.chord_l = nx / (geo.DX1 - geo.DX0)     # HTML:77
        self.mask = np.zeros((ny, nx), np.uint8)
        self.reset(u0)
        self.max_s, self.cp_min, self.cp_max = 0.6, -1.0, 1.0   # HTML:593
        self.cl_smooth = None
        self.cd_smooth = None
        self.sep_frac = 0.0
        self.me_hist = []
        self.clamp_hits = 0
        self.nsteps = 0

    def reset(self, u0=None):
        """initSim (HTML:492-500): both ping-pong sets = equilibrium at u0."""
        if u0 is not None:
            self.u0 = float(u0)
        self.F, self.rho, self.ux, self.uy = init(self.nx, self.ny, self.u0)
        self.G = self.F.copy()

    def apply_geometry(self, base_coords, a_deg):
        """applyGeometry (HTML:579-586): new mask, flow NOT re-initialised."""
        self.xp, self.yp, self.mask = geo.build_geometry(base_coords, a_deg, self.nx, self.ny)
        self.a_deg = a_deg

    def set_mask(self, mask):
        self.mask = np.ascontiguousarray(mask, dtype=np.uint8)

    def step(self, n=1):
        for _ in range(n):
            fx, fy, hits = step(self.mask, self.F, self.G, self.rho, self.ux, self.uy,
                                self.tau, self.u0)
            self.F, self.G = self.G, self.F
            self.me_hist.append((fx, fy))
            self.clamp_hits += hits
            self.nsteps += 1

    # --- diagnostics -------------------------------------------------------
    def update_fields(self, want_fields=False):
        out, U, V, Cp = field_stats(self.mask, self.rho, self.ux, self.uy, self.u0, want_fields)
        mx, c_min, c_max = out
        if mx > 0:
            self.max_s = mx
        if math.isfinite(c_min):
            self.cp_min = c_min
        if math.isfinite(c_max):
            self.cp_max = c_max
        return U, V, Cp

    def q(self):
        return 0.5 * self.u0 * self.u0 * self.chord_l      # HTML:676

    def compute_forces(self):
        """computeForces (HTML:650-700) including both EMAs."""
        fx, fy, any_, surf, rev = forces_raw(self.mask, self.rho, self.ux)
        if not any_:
            return None
        q = self.q()
        cl_raw, cd_raw = fy / q, fx / q
        self.cl_smooth = cl_raw if self.cl_smooth is None else self.cl_smooth * 0.9 + cl_raw * 0.1
        self.cd_smooth = cd_raw if self.cd_smooth is None else self.cd_smooth * 0.9 + cd_raw * 0.1
        if surf > 0:
            self.sep_frac = self.sep_frac * 0.85 + (rev / surf) * 0.15
        return dict(fx=fx, fy=fy, CL_raw=cl_raw, CD_raw=cd_raw, surf=int(surf), rev=int(rev))

    def me_coeffs(self, k=-1):
        """Momentum-exchange CL/CD of step k (not in the reference; see lbm_ref.c)."""
        fx, fy = self.me_hist[k]
        q = self.q()
        return (fy / ME_SCALE) / q, (fx / ME_SCALE) / q

    def reynolds(self):
        nu_l = (self.tau - 0.5) / 3                        # HTML:79
        return self.u0 * self.chord_l / nu_l               # HTML:865

    def stall_state(self):
        """HTML:869-884."""
        x = self.sep_frac * 100
        sep_pct = math.floor(x) + (1 if x - math.floor(x) >= 0.5 else 0)    # Math.round (ties towards +inf)
        if sep_pct < 5:
            return "Attached", sep_pct
        if sep_pct < 25:
            return f"{sep_pct}% sep", sep_pct
        return f"STALL ≈ {sep_pct}% sep", sep_pct

    def render(self, mode):
        return render_scalar(self.mask, self.rho, self.ux, self.uy, mode, self.u0,
                             self.max_s, self.cp_min, self.cp_max, self.VORT_SCALE)

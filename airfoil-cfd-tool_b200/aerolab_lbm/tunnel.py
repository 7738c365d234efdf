"""`WindTunnel`: the reference page's control surface, driven from Python.

Mirrors the operator surface of ``pages/airfoil_flow_lbm_aerolab.html``
("HTML:n"): ``applyGeometry`` (579), the ``U0`` slider (956-959), ``TAU`` (78),
``fieldMode`` (527, 953), ``simStep`` (510), ``readMacro`` (547),
``updateFieldsFromMacro`` (596), ``computeForces`` (650), ``updateStatsUI``
(862) and ``frame`` (902); and the Python bridge ``build_lbm_component`` of
``pages/Airfoil_Analysis.py:20-42``.

Every computation happens in libaerolab_lbm.so on the GPU; this file only
marshals NumPy buffers through the C ABI.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Iterable, Optional, Sequence

import numpy as np

from . import _ffi
from . import geometry as geom
from ._ffi import AerolabLbmError, check, ptr

FIELD_MODES = {"speed": 0, "cp": 1, "vort": 2, 0: 0, 1: 1, 2: 2}

# reference constants
DEFAULT_NX, DEFAULT_NY = 320, 160     # HTML:76
DEFAULT_U0 = 0.06                     # HTML:472
DEFAULT_TAU = 0.58                    # HTML:78
DEFAULT_ALPHA = 6.0                   # HTML:26, 970
STEPS_PER_FRAME = 4                   # HTML:80
FORCES_EVERY_FRAMES = 3               # HTML:914
# backend limits reused for validation (main.py:39-45)
MIN_POINTS, MAX_POINTS = 10, 500


def state_hash_numpy(f: np.ndarray, nx_global: int, gy0: int = 0) -> np.ndarray:
    """NumPy twin of ``alb_state_hash`` for populations of shape (9, rows, nx) whose first row is
    global row ``gy0`` (tests compare the device checksum with this one)."""
    f = np.ascontiguousarray(f, dtype=np.float32)
    _, rows, nx = f.shape
    g = ((np.arange(rows, dtype=np.uint64)[:, None] + np.uint64(gy0)) * np.uint64(nx_global)
         + np.arange(nx, dtype=np.uint64)[None, :])
    out = np.zeros(9, np.uint64)
    with np.errstate(over="ignore"):
        for i in range(9):
            z = g * np.uint64(0x9E3779B97F4A7C15) + f[i].view(np.uint32).astype(np.uint64) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            out[i] = z.sum(dtype=np.uint64)
    return out


def write_png(path: str, rgba: np.ndarray) -> str:
    """RGBA8 image of shape (rows, columns, 4), row 0 = bottom of the lattice, as a PNG file.
    Pure-Python encoder (zlib), no imaging dependency."""
    import struct
    import zlib
    img = np.ascontiguousarray(rgba[::-1])               # PNG rows run top to bottom
    h, w = img.shape[:2]
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)

    png = (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0))
           + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))
    with open(path, "wb") as fh:
        fh.write(png)
    return path


class WindTunnel:
    """One D2Q9 lattice on one GPU.

    ``nx`` x ``ny`` cells over the fixed world window x in [-0.42, 1.42],
    y in [-0.46, 0.46] (HTML:73); cells are square when nx = 2*ny.
    """

    def __init__(self, nx: int = DEFAULT_NX, ny: int = DEFAULT_NY, device: int = 0,
                 u0: float = DEFAULT_U0, tau: float = DEFAULT_TAU, *,
                 y0: int = 0, ny_local: Optional[int] = None):
        self._lib = _ffi.lib()
        self._h = C.c_void_p()
        self.nx, self.ny = int(nx), int(ny)
        self.y0 = int(y0)
        self.ny_local = self.ny if ny_local is None else int(ny_local)
        self.device = int(device)
        check(self._lib.alb_create_slab(self.nx, self.ny, self.y0, self.ny_local, self.device,
                                        C.byref(self._h)))
        self.name = ""
        self.coords: Optional[np.ndarray] = None
        self.alpha = DEFAULT_ALPHA
        if u0 != DEFAULT_U0 or tau != DEFAULT_TAU:
            self.reset(u0)
            self.set_tau(tau)

    # -- lifetime ---------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.alb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, code: int) -> None:
        check(code, self._h)

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    @property
    def shape(self):
        return (self.ny_local, self.nx)

    # -- geometry ---------------------------------------------------------------
    def load_coords(self, coords: Iterable[Sequence[float]], name: str = "", alpha: Optional[float] = None):
        """Use injected coordinates (``USER_COORDS``, HTML:561-563) and rasterise them."""
        arr = np.ascontiguousarray(np.asarray(list(coords), dtype=np.float64).reshape(-1, 2))
        if arr.shape[0] < 2:
            raise AerolabLbmError(_ffi.ALB_ERR_INVALID, "need at least 2 coordinate pairs")
        self.coords = arr
        self.name = name
        return self.set_alpha(self.alpha if alpha is None else alpha)

    def load_shape(self, key: str, alpha: Optional[float] = None):
        """One of the page's built-in ``SHAPES`` (HTML:123-129)."""
        return self.load_coords(geom.SHAPES[key](), name=key, alpha=alpha)

    def load_naca(self, digits: str, alpha: Optional[float] = None):
        return self.load_coords(geom.naca_digits(digits), name=f"NACA {digits}", alpha=alpha)

    def load_dat(self, path: str, parser: Optional[Callable] = None, alpha: Optional[float] = None):
        """Parse a ``.dat`` file and inject it like the Streamlit page does.

        ``parser`` defaults to the host application's own
        ``main.parse_dat_file`` (main.py:59-113), which returns
        ``(coords, fixes)``; the coordinates are then rounded to 6 decimals as
        in pages/Airfoil_Analysis.py:34-36.
        """
        from .dat import resolve_parser
        coords, fixes = resolve_parser(parser)(path)
        self.parser_fixes = fixes
        return self.load_coords(geom.round_coords(coords), name=str(path), alpha=alpha)

    def set_alpha(self, alpha_deg: float, want_mask: bool = False):
        """``applyGeometry`` (HTML:579-586): new mask, flow NOT re-initialised."""
        if self.coords is None:
            raise AerolabLbmError(_ffi.ALB_ERR_STATE, "no geometry loaded")
        out = np.empty(self.shape, np.uint8) if want_mask else None
        self._ck(self._lib.alb_rasterize(self._h, ptr(self.coords), self.coords.shape[0],
                                         float(alpha_deg), ptr(out)))
        self.alpha = float(alpha_deg)
        return out if want_mask else self

    def set_mask(self, mask_global: np.ndarray):
        m = np.ascontiguousarray(mask_global, dtype=np.uint8)
        if m.shape != (self.ny, self.nx):
            raise AerolabLbmError(_ffi.ALB_ERR_INVALID, f"mask must be {(self.ny, self.nx)}")
        self._ck(self._lib.alb_set_mask(self._h, ptr(m)))
        return self

    def mask(self) -> np.ndarray:
        out = np.empty(self.shape, np.uint8)
        self._ck(self._lib.alb_get_mask(self._h, ptr(out)))
        return out

    def panels(self):
        xp = np.empty(_ffi.ALB_NPANEL + 1)
        yp = np.empty(_ffi.ALB_NPANEL + 1)
        self._ck(self._lib.alb_get_panels(self._h, ptr(xp), ptr(yp)))
        return xp, yp

    # -- parameters -------------------------------------------------------------
    def params(self):
        u0, tau = C.c_double(), C.c_double()
        self._ck(self._lib.alb_get_params(self._h, C.byref(u0), C.byref(tau)))
        return u0.value, tau.value

    def set_u0(self, u0: float):
        """Inlet-speed slider (HTML:956-959): changes the uniform only."""
        self._ck(self._lib.alb_set_params(self._h, float(u0), self.params()[1]))
        return self

    def set_tau(self, tau: float):
        """Relaxation time (a constant 0.58 in the reference, HTML:78)."""
        self._ck(self._lib.alb_set_params(self._h, self.params()[0], float(tau)))
        return self

    def set_viscosity(self, nu_lattice: float):
        """nu = (tau - 0.5)/3 (HTML:79)."""
        return self.set_tau(3.0 * float(nu_lattice) + 0.5)

    def reset(self, u0: Optional[float] = None):
        """``initSim`` (HTML:492-500)."""
        self._ck(self._lib.alb_reset(self._h, float(self.params()[0] if u0 is None else u0)))
        self._lib.alb_reset_force_emas(self._h)
        return self

    # -- stepping ---------------------------------------------------------------
    def step(self, n: int = 1):
        """``simStep`` x n (HTML:510-525); asynchronous."""
        self._ck(self._lib.alb_step(self._h, int(n)))
        return self

    def sync(self):
        self._ck(self._lib.alb_sync(self._h))
        return self

    @property
    def steps(self) -> int:
        v = C.c_longlong()
        self._ck(self._lib.alb_step_count(self._h, C.byref(v)))
        return v.value

    def last_step_ms(self) -> float:
        v = C.c_float()
        self._ck(self._lib.alb_last_step_ms(self._h, C.byref(v)))
        return v.value

    # -- state ------------------------------------------------------------------
    def populations(self) -> np.ndarray:
        out = np.empty((9,) + self.shape, np.float32)
        self._ck(self._lib.alb_get_populations(self._h, ptr(out)))
        return out

    def population_rows(self, row0: int, nrows: int) -> np.ndarray:
        """Rows row0 .. row0+nrows-1 of this slab's populations (0 = first owned row; -1 and
        ny_local are the ghost rows), shape (9, nrows, nx)."""
        out = np.empty((9, int(nrows), self.nx), np.float32)
        self._ck(self._lib.alb_get_population_rows(self._h, int(row0), int(nrows), ptr(out)))
        return out

    def state_hash(self) -> np.ndarray:
        """Nine uint64 words, one per population plane (``alb_state_hash``); the words of all
        slabs of a lattice add up modulo 2^64 to those of the whole lattice."""
        out = np.zeros(9, np.uint64)
        self._ck(self._lib.alb_state_hash(self._h, ptr(out)))
        return out

    def set_populations(self, f: np.ndarray):
        f = np.ascontiguousarray(f, dtype=np.float32)
        if f.shape != (9,) + self.shape:
            raise AerolabLbmError(_ffi.ALB_ERR_INVALID, f"populations must be {(9,) + self.shape}")
        self._ck(self._lib.alb_set_populations(self._h, ptr(f)))
        return self

    def macro(self):
        """``readMacro`` (HTML:547-552): rho, ux, uy of the current state."""
        rho = np.empty(self.shape, np.float32)
        ux = np.empty(self.shape, np.float32)
        uy = np.empty(self.shape, np.float32)
        self._ck(self._lib.alb_get_macro(self._h, ptr(rho), ptr(ux), ptr(uy)))
        return rho, ux, uy

    def set_macro(self, rho, ux, uy):
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (rho, ux, uy)]
        self._ck(self._lib.alb_set_macro(self._h, *[ptr(a) for a in arrs]))
        return self

    def total_mass(self) -> float:
        v = C.c_double()
        self._ck(self._lib.alb_total_mass(self._h, C.byref(v)))
        return v.value

    def dump_state(self) -> dict:
        rho, ux, uy = self.macro()
        u0, tau = self.params()
        return dict(f=self.populations(), rho=rho, ux=ux, uy=uy, mask=self.mask(), u0=u0, tau=tau,
                    steps=self.steps, alpha=self.alpha, nx=self.nx, ny=self.ny)

    def load_state(self, st: dict):
        self._ck(self._lib.alb_set_params(self._h, float(st["u0"]), float(st["tau"])))
        if self.ny_local == self.ny:
            self.set_mask(st["mask"])
        self.set_populations(st["f"])
        self.set_macro(st["rho"], st["ux"], st["uy"])
        self.alpha = float(st.get("alpha", self.alpha))
        return self

    # -- diagnostics ------------------------------------------------------------
    def update_stats(self, want_fields: bool = False):
        """``updateFieldsFromMacro`` (HTML:596-614).  Returns dict(maxS, cpMin, cpMax[, U, V, Cp])."""
        st = np.zeros(3)
        U = V = Cp = None
        if want_fields:
            U = np.empty(self.shape, np.float32)
            V = np.empty(self.shape, np.float32)
            Cp = np.empty(self.shape, np.float32)
        self._ck(self._lib.alb_update_stats(self._h, ptr(st), ptr(U), ptr(V), ptr(Cp)))
        out = dict(maxS=st[0], cpMin=st[1], cpMax=st[2])
        if want_fields:
            out.update(U=U, V=V, Cp=Cp)
        return out

    def stats(self):
        st = np.zeros(3)
        self._ck(self._lib.alb_get_stats(self._h, ptr(st)))
        return dict(maxS=st[0], cpMin=st[1], cpMax=st[2])

    def field(self, mode="speed") -> np.ndarray:
        """Scalar the render shader feeds to its palette (HTML:395-420); NaN in solids."""
        out = np.empty(self.shape, np.float32)
        self._ck(self._lib.alb_get_field(self._h, FIELD_MODES[mode], ptr(out)))
        return out

    def rgba(self, mode="speed") -> np.ndarray:
        out = np.empty(self.shape + (4,), np.uint8)
        self._ck(self._lib.alb_get_rgba(self._h, FIELD_MODES[mode], ptr(out)))
        return out

    def macro_edges(self):
        """(lo, hi): ux, uy of the first and last owned row, shape (2, nx) each -- what the
        neighbouring slabs need for their vorticity taps (HTML:411-418)."""
        lo = np.empty((2, self.nx), np.float32)
        hi = np.empty((2, self.nx), np.float32)
        self._ck(self._lib.alb_get_macro_edges(self._h, ptr(lo), ptr(hi)))
        return lo, hi

    def set_macro_ghosts(self, below=None, above=None):
        """ux, uy of the rows just outside this slab (the neighbours' ``macro_edges``), (2, nx) each."""
        b = None if below is None else np.ascontiguousarray(below, dtype=np.float32)
        a = None if above is None else np.ascontiguousarray(above, dtype=np.float32)
        self._ck(self._lib.alb_set_macro_ghosts(self._h, ptr(b), ptr(a)))
        return self

    def forces(self) -> dict:
        """``computeForces`` (HTML:650-700) plus the momentum-exchange force of the last step."""
        o = np.zeros(10)
        self._ck(self._lib.alb_compute_forces(self._h, ptr(o)))
        out = dict(fx=o[0], fy=o[1], CL_raw=o[2], CD_raw=o[3], CL=o[4], CD=o[5], sep_frac=o[6],
                   surf=int(o[7]), rev=int(o[8]), any=bool(o[9]))
        if self.steps > 0:
            m = np.zeros(4)
            self._ck(self._lib.alb_get_me_forces(self._h, ptr(m)))
            out.update(Fx_me=m[0], Fy_me=m[1], CL_me=m[2], CD_me=m[3])
        return out

    def forces_partial(self) -> np.ndarray:
        o = np.zeros(4)
        self._ck(self._lib.alb_forces_partial(self._h, ptr(o)))
        return o

    def me_history(self, n: int) -> np.ndarray:
        """Fixed-point (2^-40) momentum-exchange Fx, Fy of the last n steps, shape (n, 2) int64."""
        out = np.zeros((int(n), 2), np.int64)
        self._ck(self._lib.alb_get_me_history(self._h, int(n), ptr(out)))
        return out

    def clamp_hits(self) -> int:
        v = C.c_longlong()
        self._ck(self._lib.alb_clamp_hits(self._h, C.byref(v)))
        return v.value

    def reynolds(self) -> float:
        v = C.c_double()
        self._ck(self._lib.alb_reynolds(self._h, C.byref(v)))
        return v.value

    def stall_state(self) -> str:
        """Text of the separation card (HTML:869-884)."""
        st, pct = C.c_int(), C.c_int()
        self._ck(self._lib.alb_stall_state(self._h, C.byref(st), C.byref(pct)))
        if st.value == 0:
            return "Attached"
        if st.value == 1:
            return f"{pct.value}% sep"
        return f"STALL ≈ {pct.value}% sep"

    # -- tracer particles (HTML:721-808) -------------------------------------------
    def init_particles(self, n: int = 2600, seed: int = 0):
        """``initParts`` with NPART = n (page default 2600)."""
        self._ck(self._lib.alb_particles_init(self._h, int(n), int(seed)))
        return self

    def resize_particles(self, n: int):
        """The trail-count slider (HTML:961-967)."""
        self._ck(self._lib.alb_particles_resize(self._h, int(n)))
        return self

    def step_particles(self, dt_ms: float = 16.0):
        """``stepParticles(dt)`` on the current macroscopic fields."""
        self._ck(self._lib.alb_particles_step(self._h, float(dt_ms)))
        return self

    def particles(self) -> np.ndarray:
        """(n, 8) float64: x, y, life, lane, x0, y0, speed/U0, respawned."""
        n = C.c_int()
        self._ck(self._lib.alb_particles_get(self._h, None, C.byref(n)))
        out = np.zeros((n.value, 8))
        if n.value:
            self._ck(self._lib.alb_particles_get(self._h, ptr(out), C.byref(n)))
        return out

    # -- multi-GPU y-slabs (one-row population halo) -----------------------------
    def connect_local(self, lo: "Optional[WindTunnel]", hi: "Optional[WindTunnel]"):
        """Neighbouring slabs driven by the same process (lo = below, hi = above)."""
        self._ck(self._lib.alb_connect_local(self._h, lo._h if lo is not None else None,
                                             hi._h if hi is not None else None))
        return self

    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(_ffi.ALB_IPC_BYTES)
        self._ck(self._lib.alb_ipc_export(self._h, buf))
        return buf.raw

    def ipc_connect(self, lo_blob: Optional[bytes], hi_blob: Optional[bytes]):
        lo = C.create_string_buffer(lo_blob, _ffi.ALB_IPC_BYTES) if lo_blob else None
        hi = C.create_string_buffer(hi_blob, _ffi.ALB_IPC_BYTES) if hi_blob else None
        self._ck(self._lib.alb_ipc_connect(self._h, lo, hi))
        return self

    def halo_prime(self):
        self._ck(self._lib.alb_halo_prime(self._h))
        return self

    def set_external_halo(self, on: bool):
        self._ck(self._lib.alb_set_external_halo(self._h, int(bool(on))))
        return self

    def selftest_division(self, pairs: int = 1 << 30, seed: int = 1) -> dict:
        """Compare the kernels' shared-reciprocal division with IEEE division on generated operands."""
        out = np.zeros(3, dtype=np.uint64)
        self._ck(self._lib.alb_selftest_division(self._h, int(seed), int(pairs), ptr(out)))
        return {"checked": int(out[0]), "accepted": int(out[1]), "wrong": int(out[2])}

    def launch_count(self) -> int:
        """CUDA kernels launched by this tunnel's step batches so far (graph replays included)."""
        n = C.c_longlong(0)
        self._ck(self._lib.alb_launch_count(self._h, C.byref(n)))
        return n.value

    def double_steps_active(self) -> bool:
        a = C.c_int(0)
        self._ck(self._lib.alb_get_double_steps(self._h, None, C.byref(a)))
        return bool(a.value)

    def div_mode(self) -> int:
        """0: the verified three-instruction division by tau is in use; 1: IEEE division."""
        m = C.c_int(0)
        self._ck(self._lib.alb_get_div_mode(self._h, C.byref(m)))
        return m.value

    def set_div_mode(self, mode: int):
        """1 forces IEEE division by tau, -1 returns to automatic (bit-identical results either way)."""
        self._ck(self._lib.alb_set_div_mode(self._h, int(mode)))
        return self

    def step2_plan(self, nsm: int = 148) -> dict:
        """Tiling of the fused two-step kernel for this slab (``alb_debug_step2_plan``)."""
        out = (C.c_int * 5)()
        check(self._lib.alb_debug_step2_plan(self.nx, self.ny_local, int(nsm), out))
        return dict(nseg=out[0], wo=out[1], hs=out[2], nunits=out[3], warps=out[4])

    def set_double_steps(self, mode: int):
        """-1 automatic, 0 never, 1 always: two steps per pass over HBM (bit-identical results)."""
        self._ck(self._lib.alb_set_double_steps(self._h, int(mode)))
        return self

    def halo_ptrs(self):
        """Device addresses (ints) of the rows crossing each face in the CURRENT state:
        dict(send_lo, send_hi, recv_lo, recv_hi), three pointers each, nx floats per row."""
        arrs = [(C.c_void_p * 3)() for _ in range(4)]
        self._ck(self._lib.alb_halo_ptrs(self._h, *arrs))
        keys = ("send_lo", "send_hi", "recv_lo", "recv_hi")
        return {k: [int(a[i] or 0) for i in range(3)] for k, a in zip(keys, arrs)}

    def stats_partial(self) -> np.ndarray:
        o = np.zeros(3)
        self._ck(self._lib.alb_stats_partial(self._h, ptr(o), None, None, None))
        return o

    def set_stats(self, max_s: float, cp_min: float, cp_max: float):
        self._ck(self._lib.alb_set_stats(self._h, float(max_s), float(cp_min), float(cp_max)))
        return self

    def set_params(self, u0: float, tau: float):
        self._ck(self._lib.alb_set_params(self._h, float(u0), float(tau)))
        return self

    # -- the reference's frame loop ---------------------------------------------
    FRAME_COLUMNS = ("CL", "CD", "sep_frac", "CL_raw", "CD_raw", "surf", "rev", "maxS", "cpMin", "cpMax",
                     "CL_me", "CD_me")

    def run_frames(self, nframes: int, controls=None, steps_per_frame: int = STEPS_PER_FRAME,
                   forces_every: int = FORCES_EVERY_FRAMES) -> dict:
        """``frame()`` x nframes (HTML:902-930) without host synchronisation inside the loop.

        ``controls``: optional (nframes, 2) array of (U0, tau) per frame (the sliders).  Returns a
        dict of per-frame arrays, see ``FRAME_COLUMNS``: EMA-smoothed CL/CD and the separation
        fraction (updated every ``forces_every``-th frame, as on the page), raw coefficients (NaN
        on the other frames), autoscale values, momentum-exchange coefficients."""
        ctrl = None
        if controls is not None:
            ctrl = np.ascontiguousarray(controls, dtype=np.float64).reshape(int(nframes), 2)
        series = np.empty((int(nframes), _ffi.ALB_FRAME_ROW))
        self._ck(self._lib.alb_run_frames(self._h, int(nframes), int(steps_per_frame), int(forces_every),
                                          ptr(ctrl), ptr(series)))
        return {k: series[:, i] for i, k in enumerate(self.FRAME_COLUMNS)}

    def frames_enqueue(self, nframes: int, controls=None, steps_per_frame: int = STEPS_PER_FRAME,
                       forces_every: int = FORCES_EVERY_FRAMES):
        """First half of ``run_frames``: enqueue without waiting (pair with ``frames_collect``)."""
        ctrl = None
        if controls is not None:
            ctrl = np.ascontiguousarray(controls, dtype=np.float64).reshape(int(nframes), 2)
        self._ck(self._lib.alb_frames_enqueue(self._h, int(nframes), int(steps_per_frame), int(forces_every),
                                              ptr(ctrl)))
        self._pending_frames = int(nframes)
        return self

    def frames_collect_raw(self) -> np.ndarray:
        """Second half of ``run_frames`` as the raw (nframes, 12) record array.  On a slab of a
        decomposed lattice the records are partial reductions (include/aerolab_lbm.h)."""
        n = getattr(self, "_pending_frames", 0)
        series = np.empty((n, _ffi.ALB_FRAME_ROW))
        self._ck(self._lib.alb_frames_collect(self._h, ptr(series) if n else None))
        self._pending_frames = 0
        return series

    def frames_collect(self) -> dict:
        n = getattr(self, "_pending_frames", 0)
        series = np.empty((n, _ffi.ALB_FRAME_ROW))
        self._ck(self._lib.alb_frames_collect(self._h, ptr(series) if n else None))
        self._pending_frames = 0
        return {k: series[:, i] for i, k in enumerate(self.FRAME_COLUMNS)}

    def frame(self, want_field: Optional[str] = None) -> dict:
        """One animation frame (HTML:902-930): 4 steps, render with the previous frame's
        autoscale, refresh the autoscale, forces every 3rd frame."""
        prev = self.stats()
        s = self.run_frames(1)
        out = {"stats": dict(maxS=s["maxS"][0], cpMin=s["cpMin"][0], cpMax=s["cpMax"][0])}
        if want_field is not None:
            # the page renders BEFORE updateFieldsFromMacro (HTML:909-911)
            self.set_stats(prev["maxS"], prev["cpMin"], prev["cpMax"])
            out["field"] = self.field(want_field)
            self.set_stats(**{k: out["stats"][v] for k, v in (("max_s", "maxS"), ("cp_min", "cpMin"), ("cp_max", "cpMax"))})
        if not np.isnan(s["surf"][0]):
            out["forces"] = dict(CL=s["CL"][0], CD=s["CD"][0], CL_raw=s["CL_raw"][0], CD_raw=s["CD_raw"][0],
                                 sep_frac=s["sep_frac"][0], surf=int(s["surf"][0]), rev=int(s["rev"][0]),
                                 any=bool(s["surf"][0] > 0), CL_me=s["CL_me"][0], CD_me=s["CD_me"][0])
        return out

    def save_png(self, path: Optional[str] = None, mode="speed") -> str:
        """Write the colour-mapped field (page palettes, HTML:371-393) as a PNG; row 0 of the
        lattice is the bottom of the image.  Default file name as in the page's export
        (HTML:990-992).  Pure-Python encoder (zlib), no imaging dependency."""
        return write_png(path or self.png_name(), self.rgba(mode))

    def png_name(self) -> str:
        """File name convention of the page's PNG export (HTML:990-992)."""
        base = (self.name or "airfoil").split()
        return f"{'_'.join(base)}_alpha{self.alpha:.1f}deg_lbm.png"


class LocalMultiTunnel:
    """One process, several GPUs: y-slabs created and connected by ``alb_create_multi``.

    Same control surface as ``WindTunnel`` for the calls that make sense on a decomposed
    lattice; whole-lattice arrays are assembled on the host on request."""

    def __init__(self, nx: int, ny: int, devices: Sequence[int], u0: float = DEFAULT_U0, tau: float = DEFAULT_TAU):
        lib = _ffi.lib()
        n = len(devices)
        devs = (C.c_int * n)(*[int(d) for d in devices])
        handles = (C.c_void_p * n)()
        check(lib.alb_create_multi(int(nx), int(ny), devs, n, handles))
        self._lib, self.nx, self.ny = lib, int(nx), int(ny)
        self._handles = handles
        self.slabs = []
        y0 = 0
        for k in range(n):
            t = WindTunnel.__new__(WindTunnel)          # adopt the handle created by the library
            t._lib, t._h = lib, C.c_void_p(handles[k])
            nyl = C.c_int()
            y0c = C.c_int()
            check(lib.alb_get_dims(t._h, None, None, C.byref(y0c), C.byref(nyl)), t._h)
            t.nx, t.ny, t.y0, t.ny_local, t.device = self.nx, self.ny, y0c.value, nyl.value, int(devices[k])
            t.name, t.coords, t.alpha = "", None, DEFAULT_ALPHA
            self.slabs.append(t)
        if u0 != DEFAULT_U0 or tau != DEFAULT_TAU:
            for t in self.slabs:
                t.reset(u0)
                t.set_tau(tau)

    def load_coords(self, coords, name="", alpha=None):
        for t in self.slabs:
            t.load_coords(coords, name=name, alpha=alpha)
        return self

    def load_shape(self, key, alpha=None):
        return self.load_coords(geom.SHAPES[key](), name=key, alpha=alpha)

    def set_alpha(self, alpha):
        for t in self.slabs:
            t.set_alpha(alpha)
        return self

    def set_params(self, u0, tau):
        for t in self.slabs:
            t.set_params(u0, tau)
        return self

    def reset(self, u0: Optional[float] = None):
        """``initSim`` on every slab: all slabs quiesce first (a neighbour one step behind would still
        push halo rows into buffers that are being refilled), then all are reset."""
        for t in self.slabs:
            t.sync()
        for t in self.slabs:
            t.reset(u0)
        for t in self.slabs:
            t.sync()
        self._sticky = dict(maxS=0.6, cpMin=-1.0, cpMax=1.0, cl_smooth=0.0, cd_smooth=0.0, sep_frac=0.0, ema_valid=False)
        return self

    def step(self, n: int = 1):
        check(self._lib.alb_step_multi(self._handles, len(self.slabs), int(n)))
        return self

    def sync(self):
        for t in self.slabs:
            t.sync()
        return self

    def populations(self):
        return np.concatenate([t.populations() for t in self.slabs], axis=1)

    def macro(self):
        parts = [t.macro() for t in self.slabs]
        return tuple(np.concatenate([p[k] for p in parts], axis=0) for k in range(3))

    def mask(self):
        return np.concatenate([t.mask() for t in self.slabs], axis=0)

    def forces_raw(self) -> dict:
        """Pressure-face sums and momentum exchange of the current state, summed over the slabs."""
        part = sum(t.forces_partial() for t in self.slabs)
        u0, _ = self.slabs[0].params()
        q = 0.5 * u0 * u0 * (self.nx / (geom.DX1 - geom.DX0))
        out = dict(fx=part[0], fy=part[1], surf=int(part[2]), rev=int(part[3]))
        if part[2] > 0:
            out.update(CL_raw=part[1] / q, CD_raw=part[0] / q)
        if self.slabs[0].steps > 0:
            me = sum(t.me_history(1)[0] for t in self.slabs)
            out.update(CL_me=float(me[1]) / _ffi.ALB_ME_SCALE / q, CD_me=float(me[0]) / _ffi.ALB_ME_SCALE / q)
        return out

    def update_stats(self) -> dict:
        """``updateFieldsFromMacro`` (HTML:596-614) over all slabs; the lattice-wide sticky values are
        handed to every slab (they scale its field modes)."""
        if not hasattr(self, "_sticky"):
            self._sticky = dict(maxS=0.6, cpMin=-1.0, cpMax=1.0, cl_smooth=0.0, cd_smooth=0.0, sep_frac=0.0,
                                ema_valid=False)
        parts = np.array([t.stats_partial() for t in self.slabs])
        mx, cmin, cmax = parts[:, 0].max(), parts[:, 1].min(), parts[:, 2].max()
        if mx > 0:
            self._sticky["maxS"] = float(mx)
        if np.isfinite(cmin):
            self._sticky["cpMin"] = float(cmin)
        if np.isfinite(cmax):
            self._sticky["cpMax"] = float(cmax)
        self._push_stats()
        return {k: self._sticky[k] for k in ("maxS", "cpMin", "cpMax")}

    def _push_stats(self):
        for t in self.slabs:
            t.set_stats(self._sticky["maxS"], self._sticky["cpMin"], self._sticky["cpMax"])

    def _exchange_macro_edges(self):
        edges = [t.macro_edges() for t in self.slabs]
        for k, t in enumerate(self.slabs):
            t.set_macro_ghosts(edges[k - 1][1] if k > 0 else None, edges[k + 1][0] if k + 1 < len(self.slabs) else None)

    def field(self, mode="speed") -> np.ndarray:
        """Scalar of the render shader (HTML:395-420) for the whole lattice, slab by slab."""
        if FIELD_MODES[mode] == 2:
            self._exchange_macro_edges()
        return np.concatenate([t.field(mode) for t in self.slabs], axis=0)

    def rgba(self, mode="speed") -> np.ndarray:
        if FIELD_MODES[mode] == 2:
            self._exchange_macro_edges()
        return np.concatenate([t.rgba(mode) for t in self.slabs], axis=0)

    def run_frames(self, nframes: int, controls=None, steps_per_frame: int = STEPS_PER_FRAME,
                   forces_every: int = FORCES_EVERY_FRAMES) -> dict:
        """``WindTunnel.run_frames`` on the decomposed lattice: every slab runs its frame loop on its
        own GPU without host synchronisation; the per-frame partial reductions are combined
        afterwards (``aerolab_lbm.distributed.combine_frame_partials``)."""
        from .distributed import combine_frame_partials
        for t in self.slabs:
            t.frames_enqueue(nframes, controls=controls, steps_per_frame=steps_per_frame, forces_every=forces_every)
        parts = np.stack([t.frames_collect_raw() for t in self.slabs], axis=0)
        if not hasattr(self, "_sticky"):
            self._sticky = dict(maxS=0.6, cpMin=-1.0, cpMax=1.0, cl_smooth=0.0, cd_smooth=0.0, sep_frac=0.0,
                                ema_valid=False)      # HTML:593, 641
        series = combine_frame_partials(parts, self._sticky)
        self._push_stats()
        return series

    def close(self):
        # alb_destroy_multi drains every slab before any block is freed: neighbours store halo rows
        # and flags into it
        if self.slabs:
            check(self._lib.alb_destroy_multi(self._handles, len(self.slabs)))
            for t in self.slabs:
                t._h = C.c_void_p()          # the handles are gone: WindTunnel.close()/__del__ must not free them again
        self.slabs = []


def build_lbm_component(coords_after, airfoil_name: str = "", *, nx: int = DEFAULT_NX,
                        ny: int = DEFAULT_NY, device: int = 0) -> WindTunnel:
    """Drop-in for pages/Airfoil_Analysis.py:20-42.

    Same arguments; instead of rendering an iframe it returns a running
    ``WindTunnel`` primed exactly like the page at start-up: coordinates
    rounded to 6 decimals, U0 = 0.06, tau = 0.58, alpha = 6 degrees
    (HTML:969-970).
    """
    t = WindTunnel(nx, ny, device)
    t.load_coords(geom.round_coords(coords_after), name=airfoil_name or "Uploaded airfoil",
                  alpha=DEFAULT_ALPHA)
    return t

"""Optional FastAPI endpoint: the wind tunnel as a route beside main.py's /upload_airfoil/.

    from aerolab_lbm.server import router      # in main.py, next to the existing routes
    app.include_router(router)

Conventions mirror main.py:543-628: multipart upload of a ``.dat`` file plus Form fields,
validation failures -> HTTPException(400), solver failures -> 500, blocking work in
``anyio.to_thread.run_sync`` under an ``asyncio.Semaphore``, JSON reply ``{"success": True, ...}``.
The same limits are reused (main.py:39-45): file <= 1 MiB, 10..500 points, alpha in [-10, 20], and
the route is rate limited per client like /upload_airfoil/ (main.py:543-545, "5/minute").
The coordinates go through the application's own ``parse_dat_file`` and the 6-decimal rounding
of pages/Airfoil_Analysis.py:34-36, exactly what the browser tunnel receives.
"""
from __future__ import annotations

import asyncio
import base64
import collections
import logging
import math
import os
import shutil
import tempfile
import threading
import time
from typing import Optional

from anyio import to_thread
from fastapi import APIRouter, Form, HTTPException, Request, UploadFile

from . import geometry as geom
from ._ffi import AerolabLbmError, ALB_ERR_INVALID, device_count
from .dat import resolve_parser
from .tunnel import WindTunnel

logger = logging.getLogger(__name__)
router = APIRouter()

MAX_FILE_SIZE = 1 * 1024 * 1024     # main.py:39
MAX_POINTS = 500                    # main.py:40
MIN_POINTS = 10                     # main.py:41
MIN_ALPHA, MAX_ALPHA = -10, 20      # main.py:44-45
MIN_U0, MAX_U0 = 0.030, 0.100       # slider range, HTML:41
MIN_TAU, MAX_TAU = 0.505, 2.0
MAX_CELLS = 64 * 1024 * 1024
MAX_STEPS = 200_000
MAX_CELL_UPDATES = 2 * 10 ** 12     # nx*ny*steps per request: ~20 s of one B200
MAX_FIELD_PIXELS = 4 * 1024 * 1024  # returned RGBA image: 16 MiB before base64
RATE_LIMIT, RATE_WINDOW_S = 5, 60.0   # main.py:544 `@limiter.limit("5/minute")`
LBM_DEVICE = int(os.getenv("AEROLAB_LBM_DEVICE", "0"))

_gpu_semaphore: Optional[asyncio.Semaphore] = None


def _semaphore() -> asyncio.Semaphore:
    global _gpu_semaphore
    if _gpu_semaphore is None:
        _gpu_semaphore = asyncio.Semaphore(int(os.getenv("AEROLAB_LBM_CONCURRENCY", "3")))   # main.py:47
    return _gpu_semaphore


class _RateLimiter:
    """Sliding-window limit per client address: what slowapi's ``Limiter(key_func=get_remote_address)``
    does for main.py's routes, without the dependency."""

    def __init__(self, limit: int, window_s: float):
        self.limit, self.window = limit, window_s
        self.hits = collections.defaultdict(collections.deque)
        self.lock = threading.Lock()

    def check(self, key: str, now: Optional[float] = None) -> bool:
        now = time.monotonic() if now is None else now
        with self.lock:
            q = self.hits[key]
            while q and now - q[0] >= self.window:
                q.popleft()
            if len(q) >= self.limit:
                return False
            q.append(now)
            return True


_limiter = _RateLimiter(RATE_LIMIT, RATE_WINDOW_S)


def _json_safe(x):
    """NaN / inf are not JSON: CL and CD are NaN until a frame has seen surface faces."""
    if isinstance(x, float):
        return x if math.isfinite(x) else None
    if isinstance(x, dict):
        return {k: _json_safe(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_json_safe(v) for v in x]
    if hasattr(x, "item") and getattr(x, "shape", None) == ():
        return _json_safe(x.item())
    return x


def run_tunnel_sync(coords, name, alpha, u0, tau, nx, ny, steps, field, want_field):
    """Blocking part (runs in a worker thread; ctypes releases the GIL during the calls)."""
    with WindTunnel(nx, ny, LBM_DEVICE, u0=u0, tau=tau) as t:
        t.load_coords(coords, name=name, alpha=alpha)
        frames = steps // 4
        forces = None
        for k in range(1, frames + 1):
            t.step(4)
            if k % 3 == 0:                      # reference cadence, HTML:914
                forces = t.forces()
        t.step(steps - 4 * frames)
        stats = t.update_stats()
        if forces is None:
            forces = t.forces()
        out = {
            "coefficients": {"CL": forces["CL"], "CD": forces["CD"], "CL_raw": forces["CL_raw"],
                             "CD_raw": forces["CD_raw"], "CL_me": forces.get("CL_me"),
                             "CD_me": forces.get("CD_me")},
            "separation": {"sep_frac": forces["sep_frac"], "state": t.stall_state(),
                           "surf": forces["surf"], "rev": forces["rev"]},
            "reynolds": t.reynolds(), "stats": stats, "steps": t.steps, "ms_last_call": t.last_step_ms(),
            "clamp_hits": t.clamp_hits(), "png_name": t.png_name(),
        }
        if want_field:
            rgba = t.rgba(field)
            stride = max(1, math.ceil(math.sqrt(rgba.shape[0] * rgba.shape[1] / MAX_FIELD_PIXELS)))
            if stride > 1:                       # large lattices: every stride-th pixel, at most MAX_FIELD_PIXELS
                rgba = rgba[::stride, ::stride].copy()
            out["field_stride"] = stride
            out["field"] = {"mode": field, "shape": list(rgba.shape), "encoding": "base64/rgba8, row 0 = bottom",
                            "data": base64.b64encode(rgba.tobytes()).decode("ascii")}
        return _json_safe(out)


@router.get("/lbm/health")
async def lbm_health():
    n = device_count()
    return {"status": "healthy" if n > 0 else "degraded", "cuda_devices": n}


@router.post("/lbm/run/")
async def lbm_run(
    request: Request,
    file: UploadFile,
    alpha: float = Form(6.0),
    u0: float = Form(0.06),
    tau: float = Form(0.58),
    nx: int = Form(320),
    ny: int = Form(160),
    steps: int = Form(1000),
    field: str = Form("speed"),
    return_field: bool = Form(False),
):
    client = request.client.host if request.client else "unknown"
    if not _limiter.check(client):
        raise HTTPException(status_code=429, detail=f"Rate limit exceeded: {RATE_LIMIT} per minute")
    if not (MIN_ALPHA <= alpha <= MAX_ALPHA):
        raise HTTPException(status_code=400, detail=f"Alpha must be {MIN_ALPHA} to {MAX_ALPHA} degrees")
    if not (MIN_U0 <= u0 <= MAX_U0):
        raise HTTPException(status_code=400, detail=f"U0 must be {MIN_U0} to {MAX_U0} lattice units")
    if not (MIN_TAU <= tau <= MAX_TAU):
        raise HTTPException(status_code=400, detail=f"tau must be {MIN_TAU} to {MAX_TAU}")
    if nx < 16 or ny < 8 or nx * ny > MAX_CELLS:
        raise HTTPException(status_code=400, detail=f"Lattice must be at least 16x8 and at most {MAX_CELLS} cells")
    if not (1 <= steps <= MAX_STEPS):
        raise HTTPException(status_code=400, detail=f"steps must be 1 to {MAX_STEPS}")
    if nx * ny * steps > MAX_CELL_UPDATES:
        raise HTTPException(status_code=400, detail=f"nx*ny*steps must not exceed {MAX_CELL_UPDATES:.0e} cell updates")
    if field not in ("speed", "cp", "vort"):
        raise HTTPException(status_code=400, detail="field must be speed, cp or vort")
    if not (file.filename or "").endswith(".dat"):
        raise HTTPException(status_code=400, detail="Only .dat files accepted")

    work_dir = tempfile.mkdtemp(prefix="lbm_run_")
    try:
        content = await file.read()
        if len(content) > MAX_FILE_SIZE:
            raise HTTPException(status_code=400, detail=f"File too large (max {MAX_FILE_SIZE / (1024 * 1024)}MB)")
        raw_path = os.path.join(work_dir, "raw.dat")
        with open(raw_path, "wb") as fh:
            fh.write(content)
        try:
            raw_coords, parser_fixes = resolve_parser(None)(raw_path)
        except ValueError as e:          # the stand-alone reader; main.py's parser raises HTTPException(400) itself
            raise HTTPException(status_code=400, detail=str(e))
        if len(raw_coords) > MAX_POINTS:
            raise HTTPException(status_code=400, detail=f"Too many points (max {MAX_POINTS})")
        if len(raw_coords) < MIN_POINTS:
            raise HTTPException(status_code=400, detail=f"Too few points (min {MIN_POINTS})")
        coords = geom.round_coords(raw_coords)
        name = os.path.splitext(os.path.basename(file.filename))[0]
        async with _semaphore():
            result = await to_thread.run_sync(run_tunnel_sync, coords, name, alpha, u0, tau, nx, ny, steps,
                                              field, return_field)
        return {"success": True, "coords_after": [list(map(float, p)) for p in raw_coords],
                "num_points": len(raw_coords), "parser_fixes": parser_fixes, "alpha": alpha, **result}
    except HTTPException:
        raise
    except AerolabLbmError as e:
        logger.error(str(e))
        raise HTTPException(status_code=400 if e.code == ALB_ERR_INVALID else 500, detail=str(e))
    except Exception as e:     # parser HTTPExceptions pass through above; anything else is a 500
        logger.error(str(e))
        raise HTTPException(status_code=500, detail=str(e))
    finally:
        shutil.rmtree(work_dir, ignore_errors=True)

"""Optional FastAPI endpoint: the wind tunnel as a route beside main.py's /upload_airfoil/.

    from aerolab_lbm.server import router      # in main.py, next to the existing routes
    app.include_router(router)

Conventions mirror main.py:543-628: multipart upload of a ``.dat`` file plus Form fields,
validation failures -> HTTPException(400), solver failures -> 500, blocking work in
``anyio.to_thread.run_sync`` under an ``asyncio.Semaphore``, JSON reply ``{"success": True, ...}``.
The same limits are reused (main.py:39-45): file <= 1 MiB, <= 500 points, alpha in [-10, 20].
The coordinates go through the application's own ``parse_dat_file`` and the 6-decimal rounding
of pages/Airfoil_Analysis.py:34-36, exactly what the browser tunnel receives.
"""
from __future__ import annotations

import asyncio
import base64
import logging
import os
import shutil
import tempfile
from typing import Optional

from anyio import to_thread
from fastapi import APIRouter, Form, HTTPException, UploadFile

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
LBM_DEVICE = int(os.getenv("AEROLAB_LBM_DEVICE", "0"))

_gpu_semaphore: Optional[asyncio.Semaphore] = None


def _semaphore() -> asyncio.Semaphore:
    global _gpu_semaphore
    if _gpu_semaphore is None:
        _gpu_semaphore = asyncio.Semaphore(int(os.getenv("AEROLAB_LBM_CONCURRENCY", "3")))   # main.py:47
    return _gpu_semaphore


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
            out["field"] = {"mode": field, "shape": list(rgba.shape), "encoding": "base64/rgba8, row 0 = bottom",
                            "data": base64.b64encode(rgba.tobytes()).decode("ascii")}
        return out


@router.get("/lbm/health")
async def lbm_health():
    n = device_count()
    return {"status": "healthy" if n > 0 else "degraded", "cuda_devices": n}


@router.post("/lbm/run/")
async def lbm_run(
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
        raw_coords, parser_fixes = resolve_parser(None)(raw_path)
        if len(raw_coords) > MAX_POINTS:
            raise HTTPException(status_code=400, detail=f"Too many points (max {MAX_POINTS})")
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

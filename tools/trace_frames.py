"""Timeline (AEROLAB_LBM_TRACE) of the two double steps of a frame inside alb_run_frames at configs[3]:
the plain one (steps 4k, 4k+1) and the one that carries the statistics (steps 4k+2, 4k+3)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "airfoil-cfd-tool_b200"))
import aerolab_lbm as al  # noqa: E402

nx, ny = 32768, 16384
for step in (200, 202):
    os.environ["AEROLAB_LBM_TRACE"] = str(step)
    t = al.WindTunnel(nx, ny, 0)
    t.load_shape("naca2412", alpha=5.0)
    t.run_frames(3); t.sync()
    t0 = time.perf_counter()
    t.run_frames(60)
    dt = time.perf_counter() - t0
    print(f"trace step {step}: 60 frames {dt * 1e3:.1f} ms = {dt / 60 * 1e3:.3f} ms per frame, {nx * ny * 240 / dt / 1e9:.1f} GLUPS", flush=True)
    t.sync()
    t.close()

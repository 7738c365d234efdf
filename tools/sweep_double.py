"""Single steps vs double steps over a range of lattice sizes (one GPU, graph replay, wall clock
around alb_step + sync): where the automatic rule of alb_api.cu (double_steps_enabled) should switch.
Usage: sweep_double.py NXxNY [NXxNY ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "airfoil-cfd-tool_b200"))
import aerolab_lbm as al  # noqa: E402

for arg in sys.argv[1:]:
    nx, ny = (int(v) for v in arg.split("x"))
    out = []
    for mode in (0, 1):
        t = al.WindTunnel(nx, ny, 0)
        t.set_double_steps(mode)
        t.load_shape("naca0012", alpha=5.0)
        n = max(200, min(4000, int(2e9 / (nx * ny)) // 2 * 2))
        t.step(n); t.sync()
        best = 0.0
        for rep in range(3):
            t0 = time.perf_counter()
            t.step(n); t.sync()
            best = max(best, nx * ny * n / (time.perf_counter() - t0) / 1e9)
        out.append(best)
        t.close()
    print(f"{nx}x{ny}: single {out[0]:.1f}  double {out[1]:.1f} GLUPS  ({n} steps)", flush=True)

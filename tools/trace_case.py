"""Timeline of one double step (AEROLAB_LBM_TRACE) on a single lattice: where the list-driven passes
and the fused kernel finish relative to the fork.  Usage: trace_case.py NX NY [shape alpha]"""
import os
import sys
import time

nx, ny = int(sys.argv[1]), int(sys.argv[2])
shape = sys.argv[3] if len(sys.argv) > 3 else "naca0012"
alpha = float(sys.argv[4]) if len(sys.argv) > 4 else 5.0
os.environ.setdefault("AEROLAB_LBM_TRACE", "200")
os.environ.setdefault("AEROLAB_LBM_NO_GRAPH", "1")
os.environ.setdefault("AEROLAB_LBM_DOUBLE", "1")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "airfoil-cfd-tool_b200"))
import aerolab_lbm as al  # noqa: E402

t = al.WindTunnel(nx, ny, 0)
t.load_shape(shape, alpha=alpha)
t.step(100); t.sync()
for rep in range(3):
    t0 = time.perf_counter()
    t.step(400); t.sync()
    dt = time.perf_counter() - t0
    print(f"{nx}x{ny}: {nx * ny * 400 / dt / 1e9:.1f} GLUPS (no graph), launches {t.launch_count()}", flush=True)
t.close()

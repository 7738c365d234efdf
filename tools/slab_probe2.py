"""Per-batch timing of an isolated edge slab: does a lone double step cost more than one inside a long batch?"""
import sys
sys.path.insert(0, "airfoil-cfd-tool_b200")
import aerolab_lbm as al

t = al.WindTunnel(32768, 16384, 0, y0=0, ny_local=4012)
t.load_shape("naca2412", alpha=5.0)
t.step(20); t.sync()
for n in (2, 2, 4, 4, 6, 10, 20, 40, 100):
    t.step(n)
    ms = t.last_step_ms()
    print(f"batch of {n:3d} steps: {ms:.3f} ms total, {2*ms/n:.3f} ms per double step", flush=True)

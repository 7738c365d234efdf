"""Time isolated slabs of configs[3] on one GPU (what DistributedTunnel.rebalance measures)."""
import sys
sys.path.insert(0, "airfoil-cfd-tool_b200")
import aerolab_lbm as al

nx, ny = 32768, 16384
for y0, n, label in ((0, 4012, "edge slab of a 4-GPU run (no body)"), (4012, 4160, "interior slab (half the body)"),
                     (0, 2000, "edge slab of an 8-GPU run"), (6003, 2207, "body slab of an 8-GPU run")):
    t = al.WindTunnel(nx, ny, 0, y0=y0, ny_local=n)
    t.load_shape("naca2412", alpha=5.0)
    t.step(20); t.sync()
    t.step(100)
    ms = t.last_step_ms()
    print(f"{label}: rows {y0}..{y0+n-1}: {ms/50:.3f} ms per double step, {nx*n*100/ms/1e6:.1f} GLUPS", flush=True)
    t.close()

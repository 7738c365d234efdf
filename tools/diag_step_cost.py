"""Measure the cost of the DIAG variant of the step (statistics/force reductions in the last step of a
batch) against a plain step, and one frame of alb_run_frames, at configs[3].  Run on a B200."""
import sys, time
sys.path.insert(0, "airfoil-cfd-tool_b200")
import aerolab_lbm as al
t = al.WindTunnel(32768, 16384, 0); t.load_shape("naca2412", alpha=5.0)
t.step(20); t.sync()
t.step(100); ms_plain = t.last_step_ms() / 100
t0 = time.perf_counter()
tot = 0.0
for _ in range(40):
    t.step(1); tot += t.last_step_ms()
print("plain step ms", round(ms_plain, 4), "DIAG step ms", round(tot / 40, 4), "ratio", round(tot / 40 / ms_plain, 3))
s = time.perf_counter(); t.run_frames(25); e = time.perf_counter() - s
print("frame ms", e / 25 * 1e3, "vs 4 plain", 4 * ms_plain)

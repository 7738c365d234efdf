"""Where the wall time of ONE ensemble case goes (aerolab_lbm.ensemble.run_cases does the same calls):
creation, geometry, free run, settle-phase frame loop, read-back, close."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "airfoil-cfd-tool_b200"))
import aerolab_lbm as al  # noqa: E402

nx, ny, steps = 2048, 1024, 20000
al.WindTunnel(64, 32, 0).close()
marks = [("start", time.perf_counter())]
def mark(name):
    marks.append((name, time.perf_counter()))
t = al.WindTunnel(nx, ny, 0, u0=0.06, tau=0.58); mark("create")
t.load_coords(al.SHAPES["naca0012"](), alpha=2.0); mark("load_coords")
t.sync(); mark("sync")
done = 0
while done < 18800:
    n = min(96, 18800 - done)
    t.step(n); done += n
mark("enqueue free run")
t.sync(); mark("free run done")
t.frames_enqueue(100, steps_per_frame=12, forces_every=1); mark("frames_enqueue")
s = t.frames_collect(); mark("frames_collect")
h = t.me_history(2048); mark("me_history")
st = t.stall_state(); re = t.reynolds(); ch = t.clamp_hits(); mark("scalars")
t.close(); mark("close")
for (a, ta), (b, tb) in zip(marks, marks[1:]):
    print(f"{b:>20s} {1e3 * (tb - ta):9.2f} ms")
print(f"{'total':>20s} {1e3 * (marks[-1][1] - marks[0][1]):9.2f} ms")

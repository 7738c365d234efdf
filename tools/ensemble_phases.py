"""Wall-clock breakdown of an alpha sweep with four 2048x1024 cases sharing one GPU: handle creation,
free-running phase (CUDA-graph replay, four streams), settle phase (on-device frame loops)."""
import sys, time
sys.path.insert(0, "airfoil-cfd-tool_b200")
import aerolab_lbm as al
t0 = time.perf_counter()
ts = []
for a in (0.0, 1.0, 2.0, 3.0):
    t = al.WindTunnel(2048, 1024, 0); t.load_coords(al.SHAPES["naca0012"](), alpha=a); ts.append(t)
for t in ts: t.sync()
t1 = time.perf_counter(); print("create", t1 - t0)
done = 0
while done < 18800:
    n = min(96, 18800 - done)
    for t in ts: t.step(n)
    done += n
t2 = time.perf_counter(); print("enqueue free run", t2 - t1)
for t in ts: t.sync()
t3 = time.perf_counter(); print("free run done", t3 - t1)
for t in ts: t.frames_enqueue(100, steps_per_frame=12, forces_every=1)
t4 = time.perf_counter(); print("enqueue frames", t4 - t3)
for t in ts: t.frames_collect()
t5 = time.perf_counter(); print("frames done", t5 - t3)

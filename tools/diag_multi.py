"""Diagnosis: the two-device double-step slab test with different batch sizes (tools/trip5.sh)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "airfoil-cfd-tool_b200"))
import aerolab_lbm as al

def run(nx, ny, double, batch, n, devs=(0, 1), whole_double=None, sync_whole=False):
    whole = al.WindTunnel(nx, ny, 0)
    whole.set_double_steps(double if whole_double is None else whole_double)
    whole.load_shape("naca2412", alpha=7.0)
    a = al.WindTunnel(nx, ny, devs[0], y0=0, ny_local=120)
    b = al.WindTunnel(nx, ny, devs[1], y0=120, ny_local=ny - 120)
    for s in (a, b):
        s.set_double_steps(double)
        s.load_shape("naca2412", alpha=7.0)
    a.connect_local(None, b); b.connect_local(a, None)
    whole.step(n)
    if sync_whole:
        whole.sync()
    t0 = time.time()
    try:
        for _ in range(n // batch):
            a.step(batch); b.step(batch)
        a.sync(); b.sync()
        ok = np.array_equal(np.concatenate([a.populations(), b.populations()], 1).view(np.uint32), whole.populations().view(np.uint32))
        print(f"nx={nx} double={double} batch={batch} n={n} devs={devs} whole_double={whole_double} sync_whole={sync_whole}: bitwise={ok} in {time.time()-t0:.2f}s", flush=True)
    except al.AerolabLbmError as e:
        print(f"nx={nx} double={double} batch={batch} n={n} devs={devs} whole_double={whole_double} sync_whole={sync_whole}: ERROR {e} after {time.time()-t0:.2f}s", flush=True)

if __name__ == "__main__":
    batch = int(sys.argv[1]); n = int(sys.argv[2]); devs = (0, int(sys.argv[3]))
    wd = int(sys.argv[4]) if len(sys.argv) > 4 else None
    sw = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
    run(1400, 256, 1, batch, n, devs, wd, sw)

#!/usr/bin/env python
"""BASELINE.json configs[2]: NACA 4412 at alpha = 10 deg on a 4096x2048 lattice, 50,000 steps on one
B200, with the momentum-exchange and pressure CL/CD recorded every `--every` steps (on-device frame
loop, no host synchronisation) and the achieved GLUPS / HBM fraction.

    python examples/config2_run.py --out config2_forces.csv
"""
import argparse
import csv
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "airfoil-cfd-tool_b200"))

import aerolab_lbm as al  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=4096)
    ap.add_argument("--ny", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=50000)
    ap.add_argument("--every", type=int, default=100)
    ap.add_argument("--shape", default="naca4412")
    ap.add_argument("--alpha", type=float, default=10.0)
    ap.add_argument("--out", default="config2_forces.csv")
    a = ap.parse_args()
    t = al.WindTunnel(a.nx, a.ny, 0)
    t.load_shape(a.shape, alpha=a.alpha)
    t.step(20).sync()
    t.reset()
    nframes = a.steps // a.every
    t0 = time.perf_counter()
    s = t.run_frames(nframes, steps_per_frame=a.every, forces_every=1)
    dt = time.perf_counter() - t0
    ms = t.last_step_ms()
    with open(a.out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["step", "CL_pressure_ema", "CD_pressure_ema", "CL_pressure_raw", "CD_pressure_raw", "CL_momentum_exchange",
                    "CD_momentum_exchange", "sep_frac", "maxS", "cpMin", "cpMax"])
        for k in range(nframes):
            w.writerow([(k + 1) * a.every] + [f"{s[c][k]:.6g}" for c in ("CL", "CD", "CL_raw", "CD_raw", "CL_me", "CD_me",
                                                                      "sep_frac", "maxS", "cpMin", "cpMax")])
    lups = a.nx * a.ny * nframes * a.every
    print(json.dumps({"lattice": [a.nx, a.ny], "shape": a.shape, "alpha": a.alpha, "steps": nframes * a.every,
                      "reynolds": t.reynolds(), "gpu_ms": ms, "wall_s": dt, "glups_device": lups / (ms * 1e-3) / 1e9,
                      "glups_wall": lups / dt / 1e9, "hbm_gbs": 72 * lups / (ms * 1e-3) / 1e9,
                      "clamp_hits": t.clamp_hits(), "state": t.stall_state(),
                      "final": {c: float(s[c][-1]) for c in ("CL", "CD", "CL_me", "CD_me", "sep_frac")}, "out": a.out}))


if __name__ == "__main__":
    main()

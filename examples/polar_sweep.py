#!/usr/bin/env python
"""BASELINE.json configs[4]: ensemble alpha sweep -> CL/CD polar, one case per GPU (round-robin).

    python examples/polar_sweep.py --shape naca0012 --steps 20000 --out polar.csv
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        examples/polar_sweep.py --steps 20000 --out polar.csv
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "airfoil-cfd-tool_b200"))

import aerolab_lbm as al  # noqa: E402
from aerolab_lbm import distributed as dist_mod  # noqa: E402
from aerolab_lbm.ensemble import DEFAULT_ALPHAS, alpha_sweep, write_polar_csv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="naca0012", choices=list(al.SHAPES))
    ap.add_argument("--dat", default=None, help=".dat file (uses the application's parse_dat_file)")
    ap.add_argument("--nx", type=int, default=2048)
    ap.add_argument("--ny", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--alpha-min", type=float, default=DEFAULT_ALPHAS[0])
    ap.add_argument("--alpha-max", type=float, default=DEFAULT_ALPHAS[-1])
    ap.add_argument("--alpha-step", type=float, default=1.0)
    ap.add_argument("--out", default="polar.csv")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    comm = dist_mod.init_comm(world, rank, local)
    if a.dat:
        from aerolab_lbm.dat import resolve_parser
        coords = al.round_coords(resolve_parser(None)(a.dat)[0])
    else:
        coords = al.SHAPES[a.shape]()
    n = int(round((a.alpha_max - a.alpha_min) / a.alpha_step)) + 1
    alphas = [a.alpha_min + k * a.alpha_step for k in range(n)]
    al.WindTunnel(64, 32, local).close()      # create the CUDA context / load the kernels outside the timed region
    comm.barrier()
    t0 = time.perf_counter()
    rows = alpha_sweep(coords, alphas, comm=comm, device=local, nx=a.nx, ny=a.ny, steps=a.steps)
    comm.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        write_polar_csv(rows, a.out)
        lups = len(alphas) * a.nx * a.ny * a.steps
        print(json.dumps({"cases": len(alphas), "n_gpus": world, "lattice": [a.nx, a.ny], "steps": a.steps,
                          "seconds": dt, "aggregate_glups": lups / dt / 1e9, "out": a.out,
                          "polar": [{k: r[k] for k in ("alpha", "CL", "CD", "CL_me", "CD_me", "Status")} for r in rows]}))
    comm.shutdown()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Field image of a slab-decomposed lattice (default: BASELINE.json configs[3] on all GPUs of the box).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        examples/slab_field_png.py --mode vort --steps 2000 --stride 8 --out gpurun_out/configs3_vort.png

Every rank steps its slab, the frame loop keeps the lattice-wide autoscale values
(DistributedTunnel.run_frames), every rank renders its rows with the page's palettes
(alb_get_rgba; the vorticity taps at slab edges use the neighbours' rows), and rank 0 writes
every `stride`-th pixel as a PNG (row 0 of the lattice at the bottom, like the page's canvas).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "airfoil-cfd-tool_b200"))

from aerolab_lbm import distributed as dm  # noqa: E402
from aerolab_lbm.tunnel import write_png  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=32768)
    ap.add_argument("--ny", type=int, default=16384)
    ap.add_argument("--shape", default="naca2412")
    ap.add_argument("--alpha", type=float, default=5.0)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--mode", default="vort", choices=["speed", "cp", "vort"])
    ap.add_argument("--stride", type=int, default=8)
    ap.add_argument("--out", default="slab_field.png")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    comm = dm.init_comm(world, rank, local)
    tun = dm.DistributedTunnel(a.nx, a.ny, comm, device=local)
    tun.load_shape(a.shape, alpha=a.alpha)
    series = tun.run_frames(max(1, a.steps // 4))
    img = tun.rgba(a.mode)                                  # this rank's rows, (ny_local, nx, 4)
    first = (-tun.y0) % a.stride                            # keep global rows 0, stride, 2*stride, ...
    parts = comm.gather_arrays(np.ascontiguousarray(img[first::a.stride, ::a.stride]))
    if rank == 0:
        full = np.concatenate(parts, axis=0)
        write_png(a.out, full)
        print(f"{a.out}: {full.shape[1]}x{full.shape[0]} pixels of the {a.nx}x{a.ny} lattice on {world} GPU(s), mode {a.mode}, "
              f"{tun.steps} steps, CL={series['CL'][-1]:.4f} CD={series['CD'][-1]:.4f}, {tun.stall_state()}", flush=True)
    tun.close()
    comm.shutdown()


if __name__ == "__main__":
    main()

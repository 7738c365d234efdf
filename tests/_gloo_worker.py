"""Worker of tests/test_distributed_gloo.py: one rank of a 2-process gloo job.

Drives the PRODUCT's decomposition and halo plumbing (aerolab_lbm.distributed: slab_rows, Comm,
TorchHaloExchange) on CPU tensors; the per-slab compute stand-in is the oracle (tests may use it).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "airfoil-cfd-tool_b200"))

from aerolab_lbm.distributed import HI_POPS, LO_POPS, TorchHaloExchange, init_comm, slab_rows  # noqa: E402
from oracle import geometry as ogeo  # noqa: E402
from oracle import lbm as olbm  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    nx, ny, nsteps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    out_path = sys.argv[4]
    olbm.set_threads(2)
    comm = init_comm(world, rank, 0, backend="gloo")
    y0, n = slab_rows(ny, world, rank)
    _, _, mask_full = ogeo.build_geometry(ogeo.SHAPES["naca4412"](), 10.0, nx, ny)
    m = np.zeros((n + 2, nx), np.uint8)
    lo, hi = max(0, y0 - 1), min(ny, y0 + n + 1)
    m[lo - (y0 - 1):hi - (y0 - 1)] = mask_full[lo:hi]
    F, rho, ux, uy = olbm.init(nx, n + 2, 0.06)
    state = {"F": F, "G": F.copy()}

    def rows():
        cur = state["F"]
        t = lambda i, j: torch.from_numpy(cur[i, j])          # aliases the slab's current state
        return dict(send_lo=[t(i, 1) for i in LO_POPS], send_hi=[t(i, n) for i in HI_POPS],
                    recv_lo=[t(i, 0) for i in HI_POPS], recv_hi=[t(i, n + 1) for i in LO_POPS])

    xchg = TorchHaloExchange(comm, rows)
    me = np.zeros(2, np.int64)
    for _ in range(nsteps):
        fx, fy, _ = olbm.step(m, state["F"], state["G"], rho, ux, uy, 0.58, 0.06, ny_global=ny,
                              gy0=y0 - 1, j0=1, j1=n + 1)
        state["F"], state["G"] = state["G"], state["F"]
        xchg.exchange()
        me = np.array([fx, fy], np.int64)
    me_sum = comm.allreduce(me, "sum")
    mass = comm.allreduce(np.array([olbm.total_mass(state["F"], 1, n + 1)]), "sum")
    tmax = comm.max_float(float(rank + 1))
    mn = comm.allreduce(np.array([float(rank)]), "min")
    blobs = comm.all_gather_bytes(bytes([rank]) * 4)
    parts = comm.gather_arrays(np.ascontiguousarray(state["F"][:, 1:n + 1]))
    comm.barrier()
    if rank == 0:
        np.savez(out_path, F=np.concatenate(parts, axis=1), me=me_sum, mass=mass, tmax=tmax, mn=mn,
                 blobs=np.frombuffer(b"".join(blobs), np.uint8))
    comm.shutdown()


if __name__ == "__main__":
    main()

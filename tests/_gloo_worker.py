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


def slab_partial_record(m, rho, ux, uy, n, u0, qdyn, me_fx, me_fy, do_forces):
    """NumPy restatement of the record a slab handle publishes per frame (include/aerolab_lbm.h,
    alb_run_frames): raw partial reductions over the slab's owned rows 1..n of the padded arrays."""
    own = slice(1, n + 1)
    fluid = m[own] == 0
    r64, x64, y64 = (a[own].astype(np.float64) for a in (rho, ux, uy))
    s = np.hypot(x64 / u0, y64 / u0)
    ok = fluid & (s < 4.0)
    smax = float(s[ok].max()) if ok.any() else 0.0
    cp = (r64 - 1.0) / (1.5 * u0 * u0)
    win = fluid & (cp > -4.0) & (cp < 1.2)
    rmin = float(r64[win].min()) if win.any() else np.inf
    rmax = float(r64[win].max()) if win.any() else -np.inf
    # faces, enumerated from the fluid side: a non-solid owned cell next to a solid lattice cell
    qfix = np.rint(r64 * 2.0 ** 40).astype(np.int64)
    fx = fy = surf = rev = 0
    nxl = m.shape[1]
    for dx, dy, sx, sy in ((-1, 0, -1, 0), (1, 0, 1, 0), (0, -1, 0, -1), (0, 1, 0, 1)):
        nb = np.zeros_like(fluid)
        rows = np.arange(1, n + 1) + dy
        if dx == 0:
            nb = m[rows] != 0
        elif dx == -1:
            nb[:, 1:] = m[own][:, :-1] != 0
        else:
            nb[:, :-1] = m[own][:, 1:] != 0
        face = fluid & nb
        # solid at (x+dx, y+dy): the force on the body points from the fluid towards it
        fx += int(dx * qfix[face].sum())
        fy += int(dy * qfix[face].sum())
        surf += int(face.sum())
        rev += int((face & (x64 < 0)).sum())
    rec = np.zeros(12)
    rec[0:3] = (smax, rmin, rmax)
    rec[3:9] = np.array([fx, fy, surf, rev, me_fx, me_fy], np.int64).view(np.float64)
    rec[9:12] = (u0, qdyn, 1.0 if do_forces else 0.0)
    return rec


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
    # ---- the frame loop of a decomposed lattice: per-slab raw partial records (what alb_frames_collect
    # returns on a slab handle), ONE all-gather, then the product's combine_frame_partials ---------------
    nframes = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    series = None
    if nframes:
        from aerolab_lbm.distributed import combine_frame_partials
        u0 = 0.06
        qdyn = 0.5 * u0 * u0 * (nx / (ogeo.DX1 - ogeo.DX0))
        recs = np.zeros((nframes, 12))
        for f in range(1, nframes + 1):
            for _ in range(4):
                fx, fy, _ = olbm.step(m, state["F"], state["G"], rho, ux, uy, 0.58, u0, ny_global=ny,
                                      gy0=y0 - 1, j0=1, j1=n + 1)
                state["F"], state["G"] = state["G"], state["F"]
                xchg.exchange()
            recs[f - 1] = slab_partial_record(m, rho, ux, uy, n, u0, qdyn, fx, fy, f % 3 == 0)
            me = np.array([fx, fy], np.int64)
        parts_all = comm.allgather_array(recs)
        sticky = dict(maxS=0.6, cpMin=-1.0, cpMax=1.0, cl_smooth=0.0, cd_smooth=0.0, sep_frac=0.0, ema_valid=False)
        series = combine_frame_partials(parts_all, sticky)
    me_sum = comm.allreduce(me, "sum")
    mass = comm.allreduce(np.array([olbm.total_mass(state["F"], 1, n + 1)]), "sum")
    tmax = comm.max_float(float(rank + 1))
    mn = comm.allreduce(np.array([float(rank)]), "min")
    blobs = comm.all_gather_bytes(bytes([rank]) * 4)
    parts = comm.gather_arrays(np.ascontiguousarray(state["F"][:, 1:n + 1]))
    comm.barrier()
    if rank == 0:
        extra = {} if series is None else {"series_" + k: v for k, v in series.items()}
        np.savez(out_path, F=np.concatenate(parts, axis=1), me=me_sum, mass=mass, tmax=tmax, mn=mn,
                 blobs=np.frombuffer(b"".join(blobs), np.uint8), **extra)
    comm.shutdown()


if __name__ == "__main__":
    main()

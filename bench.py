#!/usr/bin/env python
"""Benchmark of the fused D2Q9 step (BASELINE.json metric: GLUPS + HBM roofline).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload configs[i]]

A "step" is one fused lattice-Boltzmann update of the whole lattice.  Default
workload: BASELINE.json configs[3], the 32768x16384 lattice with a rasterised
NACA 2412 at alpha = 5 deg -- the configuration the 1/2/4/8-GPU metric is quoted
on.  It fits one B200 (2 x 19.3 GB of populations), so the same lattice is used
at every N (strong scaling, y-slabs, one-row NVLink halo).  One JSON line is
printed by rank 0.

--impl reference times the reference's CPU path.  The reference has no CPU
implementation of the step (it is a WebGL shader), so this is the strict-fp32
OpenMP restatement in oracle/ ("port"), on all host threads, on a bounded
sample of the same workload (a band of rows through the airfoil).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "airfoil-cfd-tool_b200")
for _p in (ROOT, PKG_DIR):
    if _p not in sys.path:
        sys.path.insert(0, _p)

BYTES_PER_LUP = 72.0        # 9 fp32 loads + 9 fp32 stores
METRIC = "d2q9_glups"
UNIT = "GLUPS"

WORKLOADS = {
    # name: (nx, ny, shape, alpha)  -- BASELINE.md section 4
    "configs[1]": (320, 160, "naca0012", 5.0),
    "configs[2]": (4096, 2048, "naca4412", 10.0),
    "configs[3]": (32768, 16384, "naca2412", 5.0),
    "configs[4]-case": (2048, 1024, "naca0012", 5.0),
}
U0, TAU = 0.06, 0.58


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def committed_traffic(workload, kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel from
    the committed `ncu --set full` capture of this workload on ONE GPU, or None.  The entry names the
    kernel it was measured on; a line for another kernel (or another N) gets no traffic figure."""
    path = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    try:
        with open(path) as fh:
            e = json.load(fh).get(workload)
        if e and e.get("kernel") == kernel:
            return e
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            time.sleep(0.15)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2])); power.append(float(p[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=max(power))
        return out


def slab_rows(ny, world, rank):
    base, rem = divmod(ny, world)
    y0 = rank * base + min(rank, rem)
    return y0, base + (1 if rank < rem else 0)


def cpu_sample(nx, ny, shape, alpha, band_rows, steps, warmup, max_seconds=None, threads=None):
    """Time the oracle (OpenMP, all host threads unless `threads` says otherwise) on a band of rows
    centred on the airfoil."""
    import numpy as np
    from oracle import geometry as ogeo
    from oracle import lbm as olbm
    band = min(band_rows, ny)
    y0 = max(0, ny // 2 - band // 2)
    nrows = band + 2
    xp, yp = ogeo.panelise(ogeo.rotate(ogeo.SHAPES[shape](), alpha))
    # rasterise only the band (+ghost rows): reuse the oracle's scan conversion row by row
    full_rows = range(y0 - 1, y0 + band + 1)
    mask = np.zeros((nrows, nx), np.uint8)
    sub = ogeo.raster_rows(xp, yp, nx, ny, full_rows)
    mask[:, :] = sub
    F, rho, ux, uy = olbm.init(nx, nrows, U0)
    G = F.copy()
    cores = threads or os.cpu_count() or 1
    olbm.set_threads(cores)
    cells = nx * band
    for _ in range(warmup):
        olbm.step(mask, F, G, rho, ux, uy, TAU, U0, ny_global=ny, gy0=y0 - 1, j0=1, j1=band + 1)
        F, G = G, F
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        olbm.step(mask, F, G, rho, ux, uy, TAU, U0, ny_global=ny, gy0=y0 - 1, j0=1, j1=band + 1)
        F, G = G, F
        done += 1
        if max_seconds is not None and time.perf_counter() - t0 > max_seconds:
            break
    dt = time.perf_counter() - t0
    return dict(glups=cells * done / dt / 1e9, seconds=dt, steps=done, cells=cells, cores=cores,
                sample=f"rows {y0}..{y0 + band - 1} of the {nx}x{ny} lattice ({cells} cells/step), "
                       f"{done} steps, {cores} OpenMP threads")


def run_reference(args, rank, world):
    """CPU arm: rank 0 only; other ranks exit without work."""
    if rank != 0:
        return
    nx, ny, shape, alpha = WORKLOADS[args.workload]
    band = 256 if nx * 256 <= 16 * 1024 * 1024 else max(16, (16 * 1024 * 1024) // nx)
    band = min(band, ny)
    r = cpu_sample(nx, ny, shape, alpha, band, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["glups"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / r["steps"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {shape} alpha={alpha} on {nx}x{ny}, U0={U0}, tau={TAU}",
                   "note": "CPU sample of the same lattice; the reference has no CPU LBM, this is the "
                           "strict-fp32 OpenMP restatement in oracle/ (kind=port)"},
        "cpu_baseline": {"value": r["glups"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["glups"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="configs[3]", choices=list(WORKLOADS))
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="weak: configs[3] width, 2048 rows per GPU (32768 x 2048*N), fixed work per GPU")
    ap.add_argument("--balance", default="auto", choices=["auto", "equal", "measured"],
                    help="rows per GPU: equal, or re-split before the run in proportion to the measured speed of "
                         "each slab (auto: measured for strong scaling on more than one GPU with the p2p halo)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-weak", action="store_true",
                    help="skip the weak-scaling companion measurement (32768 x 2048*N) that strong-scaling runs of "
                         "configs[3] append to their line")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch

    import aerolab_lbm as al
    from aerolab_lbm import distributed as dist_mod

    if al.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: aerolab_lbm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    comm = dist_mod.init_comm(world, rank, local_rank)

    nx, ny, shape, alpha = WORKLOADS[args.workload]
    if args.scaling == "weak":
        ny = 2048 * world
    tun = dist_mod.DistributedTunnel(nx, ny, comm, device=local_rank, halo=args.halo)
    tun.load_shape(shape, alpha=alpha)
    tun.sync()
    comm.barrier()
    balance = args.balance
    if balance == "auto":
        balance = "measured" if (world > 1 and args.halo == "p2p" and args.scaling == "strong") else "equal"
    if balance == "measured" and world > 1:
        tun.rebalance()          # set-up, outside every timed region; the flow starts from rest afterwards
        tun.sync()
        comm.barrier()

    cells_global = nx * ny
    # ---- device-resident throughput (`value`) ---------------------------------
    tun.step(args.warmup)
    tun.sync()
    comm.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    comm.barrier()
    launches0 = tun.t.launch_count()
    t0 = time.perf_counter()
    tun.step(args.steps)
    ms_dev = tun.last_step_ms()          # CUDA events on the launching stream; synchronises
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    comm.barrier()
    launches = tun.t.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_dev = comm.max_float(ms_dev)
    wall_ms = comm.max_float(wall_ms)
    glups = cells_global * args.steps / (ms_dev * 1e-3) / 1e9
    # checksum of the state the timed region produced (before the e2e loops add their own steps): the
    # same `warmup + steps` steps give the same nine words at every N
    state_hash, steps_at_hash = tun.state_hash(), tun.steps

    # same-box calibration of the denominator: a plain device copy (what MEASURED_PEAKS.json holds)
    copy_here = None
    if rank == 0 and world == 1:
        try:
            a = torch.empty(1 << 30, dtype=torch.bfloat16, device=f"cuda:{local_rank}")
            b = torch.empty_like(a)
            best = 1e9
            for _ in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                b.copy_(a)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            copy_here = 2 * a.numel() * 2 / (best * 1e-3) / 1e9
            del a, b
        except Exception:
            copy_here = None

    # roofline of the dominant kernel: per launch, this rank's cells.  Algorithmic traffic is
    # 72 B per cell update.  With double steps the dominant kernel (march2_kernel)
    # performs TWO updates per cell and launch while reading and writing the state once, so the
    # algorithmic rate can exceed the DRAM peak; `traffic` (ncu, per launch) and `dram_frac` show
    # what actually crosses the HBM interface.
    peak, peak_src = measured_peak()
    cells_local = nx * tun.ny_local
    doubles = tun.t.double_steps_active() and args.halo == "p2p"
    steps_per_launch = 2 if doubles else 1
    launch_ms = ms_dev / args.steps * steps_per_launch
    alg_bytes = BYTES_PER_LUP * cells_local * steps_per_launch
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
    kernel = "alb::march2_kernel" if doubles else "alb::step_kernel<MODE_STEP>"
    # the ncu capture is of the one-GPU run of this workload: per-rank launches of an N-GPU run
    # process 1/N of the cells and get no traffic figure
    cap = committed_traffic(args.workload, kernel) if world == 1 and args.scaling == "strong" else None
    traffic = cap["dram_bytes_per_launch"] if cap else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": cap.get("source") if cap else None,
                "peak_source": peak_src, "copy_gbs_measured_in_this_run": copy_here,
                "kernel": kernel + (" (two steps per launch: reads the state once, writes it once; border/body/"
                                    "slab-edge tasks go through list-driven step_kernel passes on a second stream)"
                                    if doubles else ""),
                "steps_per_launch": steps_per_launch,
                "algorithmic_bytes_per_launch": alg_bytes,
                "launch_ms": launch_ms,
                # what bounds a kernel that performs TWO updates per cell and pass over HBM: 36 B per update
                "bound_bytes_per_lup": BYTES_PER_LUP / steps_per_launch,
                "frac_of_own_bound": achieved / steps_per_launch / peak,
                "dram_frac": (traffic / (launch_ms * 1e-3) / 1e9 / peak) if traffic else None}

    # ---- end to end through the public API (`e2e`) ----------------------------
    e2e = None
    if not args.no_e2e:
        e2e = dist_mod.bench_e2e(tun, comm, args.steps, cells_global)

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------
    # the oracle on all host threads on a band of this workload (the figure next to
    # `value`), the same band on ONE thread, and configs[0] itself (320x160, 1,000 steps)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        band = 256 if nx * 256 <= 16 * 1024 * 1024 else max(16, (16 * 1024 * 1024) // nx)
        r = cpu_sample(nx, ny, shape, alpha, min(band, ny), steps=1000, warmup=2, max_seconds=10.0)
        r1 = cpu_sample(nx, ny, shape, alpha, min(band, ny), steps=1000, warmup=1, max_seconds=6.0, threads=1)
        c0nx, c0ny, c0shape, c0alpha = WORKLOADS["configs[1]"]
        c0 = cpu_sample(c0nx, c0ny, c0shape, c0alpha, c0ny, steps=1000, warmup=2, max_seconds=20.0)
        cpu = {"value": r["glups"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "one_thread": {"value": r1["glups"], "unit": UNIT, "cores": 1, "sample": r1["sample"]},
               "configs[0]": {"value": c0["glups"], "unit": UNIT, "cores": c0["cores"], "seconds": c0["seconds"],
                              "steps": c0["steps"], "sample": c0["sample"]}}

    # ---- the same loop with the frame's image, as the page draws it (`e2e_fields`) ----------
    e2e_fields = None
    if not args.no_e2e and world == 1:
        e2e_fields = dist_mod.bench_e2e_fields(tun, cells_global)

    forces = tun.forces()
    total_steps = tun.steps
    decomposition = (f"{world} y-slab(s), one-row population halo ({args.halo}), rows per GPU {balance}: {tun.rows}")

    # ---- weak-scaling companion (north_star: "strong and weak scaling") ------------------------------
    # the default line is strong scaling; the same run also times the fixed-work-per-GPU lattice
    # (configs[3] width, 2048 rows per GPU), so both curves come from one driver invocation per N
    weak = None
    if args.scaling == "strong" and args.workload == "configs[3]" and not args.no_weak:
        tun.close()
        tun = None
        comm.barrier()
        wny = 2048 * world
        wt = dist_mod.DistributedTunnel(nx, wny, comm, device=local_rank, halo=args.halo)
        wt.load_shape(shape, alpha=alpha)
        wt.step(args.warmup)
        wt.sync()
        comm.barrier()
        wt.step(args.steps)
        wms = comm.max_float(wt.last_step_ms())
        whash = wt.state_hash()
        weak = {"value": nx * wny * args.steps / (wms * 1e-3) / 1e9, "unit": UNIT, "scaling": "weak",
                "lattice": [nx, wny], "rows_per_gpu": 2048, "ms_per_step": wms / args.steps, "steps": args.steps,
                "state_hash": ["%016x" % int(v) for v in whash]}
        wt.close()
    if rank == 0:
        line = {
            "metric": METRIC, "value": glups, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {shape} alpha={alpha} on {nx}x{ny}, U0={U0}, tau={TAU}",
                       "decomposition": decomposition,
                       "l2": "populations (2 x %.1f GB per GPU) are far larger than the 126 MB L2; no flush needed"
                             % (36.0 * cells_local / 1e9),
                       "timing": "CUDA events on the launching stream, max over ranks"},
            "wall_ms_per_step": wall_ms / args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_fields": e2e_fields, "weak_scaling": weak,
            # counted by the library (alb_launch_count): kernels launched on this rank in the timed
            # region, kernels inside replayed CUDA graphs included
            "gpu_launches": int(launches),
            "clocks": clocks,
            "check": {"CL_me": forces.get("CL_me"), "CD_me": forces.get("CD_me"),
                      "CL_pressure_raw": forces.get("CL_raw"), "CD_pressure_raw": forces.get("CD_raw"),
                      "total_steps": total_steps,
                      # position-dependent checksum of the population bit patterns (alb_state_hash), slabs
                      # added modulo 2^64: identical at every N for the same number of steps
                      "state_hash": ["%016x" % int(v) for v in state_hash], "state_hash_after_steps": steps_at_hash},
        }
        print(json.dumps(line), flush=True)
    if tun is not None:
        tun.close()
    comm.shutdown()


if __name__ == "__main__":
    main()

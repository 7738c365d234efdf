"""Multi-GPU driver: y-slab decomposition with a one-row population halo.

One process per GPU (``torchrun``); ``torch.distributed`` is plumbing only
(rendezvous, exchanging CUDA-IPC blobs, scalar reductions, barriers).  The
halo itself has two transports:

``p2p`` (default)  the step kernel stores its edge rows straight into the
                   neighbour's ghost rows through CUDA-IPC-mapped peer memory
                   (NVLink), and stream-ordered flag kernels order the steps;
                   no host involvement per step.
``nccl``           a plain ``torch.distributed`` send/recv of the six rows
                   between single steps (fallback; also what the CPU ``gloo``
                   tests exercise through :class:`TorchHaloExchange`).

The reference has nothing to mirror here (it is single-GPU WebGL); the slab
rules follow SURVEY.md section 8(e): rows are contiguous, inlet/outlet columns
stay local, the bottom/top slabs own the equilibrium rows, the pull scheme
needs f2,f5,f6 from below and f4,f7,f8 from above.
"""
from __future__ import annotations

import math
import os
import time
from typing import List, Optional, Sequence

import numpy as np

LO_POPS = (4, 7, 8)     # leave through the bottom face (e_y = -1)
HI_POPS = (2, 5, 6)     # leave through the top face (e_y = +1)


def slab_rows(ny: int, world: int, rank: int):
    """Rows [y0, y0+n) owned by `rank`: equal shares, the first ny % world ranks get one more."""
    base, rem = divmod(ny, world)
    y0 = rank * base + min(rank, rem)
    return y0, base + (1 if rank < rem else 0)


class Comm:
    """Thin wrapper over torch.distributed; a no-op for a single process."""

    def __init__(self, world: int = 1, rank: int = 0, device: Optional[str] = None):
        self.world, self.rank = world, rank
        self.device = device
        self._dist = None
        if world > 1:
            import torch.distributed as dist
            self._dist = dist

    def _tensor(self, data, dtype):
        import torch
        return torch.tensor(data, dtype=dtype, device=self.device or "cpu")

    def barrier(self):
        if self._dist:
            self._dist.barrier()

    def max_float(self, x: float) -> float:
        if not self._dist:
            return float(x)
        import torch
        t = self._tensor([x], torch.float64)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        return float(t.item())

    def allreduce(self, arr: np.ndarray, op: str = "sum") -> np.ndarray:
        """Element-wise reduction of a small float64 or int64 array (int sums are exact)."""
        if not self._dist:
            return np.array(arr, copy=True)
        import torch
        dt = torch.int64 if np.issubdtype(np.asarray(arr).dtype, np.integer) else torch.float64
        t = self._tensor(np.asarray(arr).tolist(), dt)
        ops = {"sum": self._dist.ReduceOp.SUM, "max": self._dist.ReduceOp.MAX, "min": self._dist.ReduceOp.MIN}
        self._dist.all_reduce(t, op=ops[op])
        return t.cpu().numpy()

    def allgather_array(self, a: np.ndarray) -> np.ndarray:
        """Stack one equally shaped float64 array per rank along a new first axis (bit patterns are
        preserved: the payload may be integers stored in the double slots)."""
        a = np.ascontiguousarray(a, dtype=np.float64)
        if not self._dist:
            return a[None].copy()
        import torch
        t = torch.from_numpy(a.view(np.int64).copy()).to(self.device or "cpu")
        out = [torch.empty_like(t) for _ in range(self.world)]
        self._dist.all_gather(out, t)
        return np.stack([o.cpu().numpy().view(np.float64) for o in out], axis=0)

    def allgather_float(self, x: float) -> List[float]:
        """One float per rank, known to every rank afterwards."""
        v = np.zeros(self.world)
        v[self.rank] = float(x)
        return [float(a) for a in self.allreduce(v, "sum")]

    def all_gather_bytes(self, b: bytes) -> List[bytes]:
        if not self._dist:
            return [b]
        out = [None] * self.world
        self._dist.all_gather_object(out, b)
        return out

    def gather_arrays(self, a: np.ndarray, dst: int = 0):
        """Gather NumPy arrays on `dst` (slow path, for tests and dumps)."""
        if not self._dist:
            return [a]
        out = [None] * self.world
        self._dist.all_gather_object(out, a)
        return out if self.rank == dst else None

    def shutdown(self):
        if self._dist and self._dist.is_initialized():
            self._dist.destroy_process_group()


def init_comm(world: int, rank: int, local_rank: int = 0, backend: Optional[str] = None) -> Comm:
    """Initialise torch.distributed from the torchrun environment (MASTER_ADDR/PORT)."""
    if world <= 1:
        return Comm()
    import torch
    import torch.distributed as dist
    cuda = torch.cuda.is_available()
    backend = backend or ("nccl" if cuda else "gloo")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if not dist.is_initialized():
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device(f"cuda:{local_rank}")
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return Comm(world, rank, f"cuda:{local_rank}" if backend == "nccl" else "cpu")


class TorchHaloExchange:
    """Move the six halo rows with torch.distributed point-to-point calls.

    ``rows()`` must return ``dict(send_lo, send_hi, recv_lo, recv_hi)`` of three
    1-D torch tensors each (CPU tensors under gloo, CUDA tensors under NCCL) that
    alias the slab's CURRENT state.
    """

    def __init__(self, comm: Comm, rows):
        self.comm = comm
        self.rows = rows

    def exchange(self):
        dist = self.comm._dist
        if dist is None:
            return
        r = self.rows()
        rank, world = self.comm.rank, self.comm.world
        ops = []
        for k in range(3):
            if rank > 0:
                ops.append(dist.P2POp(dist.isend, r["send_lo"][k], rank - 1, tag=k))
                ops.append(dist.P2POp(dist.irecv, r["recv_lo"][k], rank - 1, tag=3 + k))
            if rank < world - 1:
                ops.append(dist.P2POp(dist.isend, r["send_hi"][k], rank + 1, tag=3 + k))
                ops.append(dist.P2POp(dist.irecv, r["recv_hi"][k], rank + 1, tag=k))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()


class _DevRow:
    """Expose a raw device address as a CUDA array so torch can wrap it without a copy."""

    def __init__(self, addr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (addr, False), "version": 2}


FRAME_COLUMNS = ("CL", "CD", "sep_frac", "CL_raw", "CD_raw", "surf", "rev", "maxS", "cpMin", "cpMax", "CL_me", "CD_me")


def combine_frame_partials(parts: np.ndarray, sticky: dict) -> dict:
    """The page's per-frame host logic (HTML:611-613 autoscale, 672-679 + 699 force EMAs) applied to
    the raw partial reductions that the slabs of a decomposed lattice record per frame
    (``alb_frames_collect`` on a slab handle; layout in include/aerolab_lbm.h).

    ``parts``: (nslabs, nframes, 12).  ``sticky``: dict(maxS, cpMin, cpMax, cl_smooth, cd_smooth,
    sep_frac, ema_valid), updated in place.  Returns per-frame arrays keyed like
    ``WindTunnel.run_frames``.  Extrema combine by max / min and the face sums are exact integers,
    and the float64 operations below are the ones ``frame_finalize_kernel`` performs on one GPU, so
    the series is identical to that of the undivided lattice."""
    parts = np.ascontiguousarray(parts, dtype=np.float64)
    nslabs, nframes, _ = parts.shape
    ints = np.ascontiguousarray(parts[:, :, 3:9]).view(np.int64)
    me_scale = float(2 ** 40)
    out = np.empty((nframes, 12))
    for f in range(nframes):
        smax = float(parts[:, f, 0].max())
        rho_min = float(parts[:, f, 1].min())
        rho_max = float(parts[:, f, 2].max())
        fx, fy, surf_i, rev_i, mfx_i, mfy_i = (int(v) for v in ints[:, f, :].sum(axis=0))
        u0, q, do_forces = float(parts[0, f, 9]), float(parts[0, f, 10]), parts[0, f, 11] != 0
        cden = (1.5 * u0) * u0
        if smax > 0:
            sticky["maxS"] = smax
        if rho_min <= rho_max:
            with np.errstate(divide="ignore", invalid="ignore"):
                cmin = float(np.float64(rho_min - 1.0) / np.float64(cden))
                cmax = float(np.float64(rho_max - 1.0) / np.float64(cden))
            if math.isfinite(cmin):
                sticky["cpMin"] = cmin
            if math.isfinite(cmax):
                sticky["cpMax"] = cmax
        cl_raw = cd_raw = surf = rev = math.nan
        if do_forces:
            surf, rev = float(surf_i), float(rev_i)
            if surf_i > 0:
                cl_raw = (float(fy) / me_scale / 3.0) / q
                cd_raw = (float(fx) / me_scale / 3.0) / q
                if not sticky["ema_valid"]:
                    sticky["cl_smooth"], sticky["cd_smooth"], sticky["ema_valid"] = cl_raw, cd_raw, True
                else:
                    sticky["cl_smooth"] = sticky["cl_smooth"] * 0.9 + cl_raw * 0.1
                    sticky["cd_smooth"] = sticky["cd_smooth"] * 0.9 + cd_raw * 0.1
                sticky["sep_frac"] = sticky["sep_frac"] * 0.85 + (rev / surf) * 0.15
        with np.errstate(divide="ignore", invalid="ignore"):
            cl_me = float(np.float64(float(mfy_i) / me_scale) / np.float64(q))
            cd_me = float(np.float64(float(mfx_i) / me_scale) / np.float64(q))
        out[f] = (sticky["cl_smooth"] if sticky["ema_valid"] else math.nan,
                  sticky["cd_smooth"] if sticky["ema_valid"] else math.nan,
                  sticky["sep_frac"], cl_raw, cd_raw, surf, rev,
                  sticky["maxS"], sticky["cpMin"], sticky["cpMax"], cl_me, cd_me)
    return {k: out[:, i] for i, k in enumerate(FRAME_COLUMNS)}


def resplit_rows(rows: Sequence[int], times: Sequence[float], ny: int, min_rows: int = 8) -> List[int]:
    """New rows per slab from the measured time of each slab: proportional to the measured speed
    (rows per millisecond), damped by averaging with the old split (the cost is not uniform inside
    a slab, so one proportional step overshoots), at least ``min_rows`` each, adding up to ``ny``."""
    speed = [r / max(t, 1e-9) for r, t in zip(rows, times)]
    tot = sum(speed)
    new = [max(min_rows, int(round(ny * sp / tot))) for sp in speed]
    out = [max(min_rows, (a + b) // 2) for a, b in zip(rows, new)]
    out[out.index(max(out))] += ny - sum(out)
    if min(out) < 1:
        raise ValueError("cannot split %d rows over %d slabs with at least %d rows each" % (ny, len(rows), min_rows))
    return out


class DistributedTunnel:
    """A lattice split into one y-slab per rank; same control surface as WindTunnel."""

    def __init__(self, nx: int, ny: int, comm: Optional[Comm] = None, device: int = 0, halo: str = "p2p",
                 u0: float = 0.06, tau: float = 0.58, rows: Optional[Sequence[int]] = None):
        """``rows``: rows per rank (bottom slab first); default: as equal as possible."""
        self.comm = comm or Comm()
        self.nx, self.ny = nx, ny
        self.device = device
        self._u0, self._tau = u0, tau
        self.halo = halo if self.comm.world > 1 else "none"
        self.cl_smooth = None
        self.cd_smooth = None
        self.sep_frac = 0.0
        self.max_s, self.cp_min, self.cp_max = 0.6, -1.0, 1.0
        self.t = None
        self._build(rows, connect=True)

    def _build(self, rows, connect: bool):
        """(Re-)create this rank's slab; ``connect=False`` leaves it isolated (timing runs)."""
        from .tunnel import WindTunnel
        if self.t is not None:
            self.t.sync()
            self.t.close()
        w, r = self.comm.world, self.comm.rank
        if rows is None:
            self.rows = [slab_rows(self.ny, w, k)[1] for k in range(w)]
        else:
            self.rows = [int(v) for v in rows]
            if len(self.rows) != w or sum(self.rows) != self.ny or min(self.rows) < 1:
                raise ValueError("rows must give every rank at least one row and add up to ny")
        self.y0, self.ny_local = sum(self.rows[:r]), self.rows[r]
        if self.ny_local < 1:
            raise ValueError("more ranks than lattice rows")
        self.t = WindTunnel(self.nx, self.ny, self.device, u0=self._u0, tau=self._tau, y0=self.y0,
                            ny_local=self.ny_local)
        self._xchg = None
        self._row_cache = {}
        if not connect:
            return
        if self.halo == "p2p":
            blobs = self.comm.all_gather_bytes(self.t.ipc_export())
            self.t.ipc_connect(blobs[r - 1] if r > 0 else None, blobs[r + 1] if r < w - 1 else None)
            self.comm.barrier()
        elif self.halo == "nccl":
            self.t.set_external_halo(True)
            self._xchg = TorchHaloExchange(self.comm, self._torch_rows)

    def rebalance(self, calib_steps: int = 12, rounds: int = 2, min_rows: int = 8):
        """Static load balancing of the slabs, before the run starts (the flow is reset).

        Rows are not equally expensive: slabs that contain the body spend extra time in the
        general-task kernels, and with equal slabs everybody waits for them every step.  Each round
        times ``calib_steps`` steps of every slab in isolation (unconnected, so no slab waits for
        another), all-gathers the times and re-splits the rows in proportion to the measured speed.
        Returns the final rows per rank.  No-op on one rank."""
        if self.comm.world == 1 or self.t.coords is None:
            return list(self.rows)
        coords, name, alpha = self.t.coords, self.t.name, self.t.alpha
        rows = list(self.rows)
        for _ in range(rounds):
            self._build(rows, connect=False)
            self.t.load_coords(coords, name=name, alpha=alpha)
            self.t.step(calib_steps)            # warm-up: graph capture, first-touch
            self.t.sync()
            self.t.step(calib_steps)
            ms = self.t.last_step_ms()
            times = self.comm.allgather_float(ms)
            rows = resplit_rows(rows, times, self.ny, min_rows)
            self.calib_ms = times
        self._build(rows, connect=True)
        self.t.load_coords(coords, name=name, alpha=alpha)
        self.comm.barrier()
        return list(self.rows)

    # -- halo rows as torch tensors (nccl transport) ---------------------------------
    def _torch_rows(self):
        import torch
        ptrs = self.t.halo_ptrs()
        key = tuple(ptrs["send_lo"])
        if key not in self._row_cache:
            dev = f"cuda:{self.device}"
            self._row_cache[key] = {k: [torch.as_tensor(_DevRow(a, self.nx), device=dev) for a in v]
                                    for k, v in ptrs.items()}
        return self._row_cache[key]

    # -- control surface ---------------------------------------------------------------
    def load_coords(self, coords, name="", alpha=None):
        self.t.load_coords(coords, name=name, alpha=alpha)
        return self

    def load_shape(self, key, alpha=None):
        self.t.load_shape(key, alpha=alpha)
        return self

    def set_alpha(self, alpha):
        self.t.set_alpha(alpha)
        return self

    def set_params(self, u0, tau):
        self.t.set_params(u0, tau)
        return self

    def reset(self, u0=None):
        """``initSim`` on every slab.  A slab may be one step behind its neighbours, and that step
        still pushes halo rows into this slab's buffers: every rank finishes its own work, all ranks
        meet, and only then are the buffers (ghost rows included) refilled."""
        self.t.sync()
        self.comm.barrier()
        self.t.reset(u0)
        self.t.sync()
        self.comm.barrier()
        self.cl_smooth = self.cd_smooth = None
        self.sep_frac = 0.0
        return self

    def step(self, n: int = 1):
        if self.halo == "nccl":
            import torch
            t0 = time.perf_counter()
            for _ in range(n):
                self.t.step(1)
                self.t.sync()
                self._xchg.exchange()
                torch.cuda.current_stream().synchronize()
            self._nccl_ms = (time.perf_counter() - t0) * 1e3     # host-synchronous transport: wall clock
        else:
            self.t.step(n)
        return self

    def sync(self):
        self.t.sync()
        return self

    def last_step_ms(self) -> float:
        """GPU time of the last step() call (CUDA events); wall clock for the host-driven nccl transport."""
        if self.halo == "nccl":
            return self._nccl_ms
        return self.t.last_step_ms()

    @property
    def steps(self) -> int:
        return self.t.steps

    def close(self):
        self.t.sync()
        self.comm.barrier()
        self.t.close()

    # -- diagnostics (partials reduced over ranks; host logic of HTML:611-613, 672-699) ----
    def update_stats(self) -> dict:
        p = self.t.stats_partial()
        mx = self.comm.allreduce(np.array([p[0], p[2]]), "max")
        mn = self.comm.allreduce(np.array([p[1]]), "min")
        if mx[0] > 0:
            self.max_s = float(mx[0])
        if math.isfinite(mn[0]):
            self.cp_min = float(mn[0])
        if math.isfinite(mx[1]):
            self.cp_max = float(mx[1])
        self.t.set_stats(self.max_s, self.cp_min, self.cp_max)
        return dict(maxS=self.max_s, cpMin=self.cp_min, cpMax=self.cp_max)

    def forces(self) -> dict:
        part = self.comm.allreduce(self.t.forces_partial(), "sum")
        fx, fy, surf, rev = (float(v) for v in part)
        u0, _ = self.t.params()
        q = 0.5 * u0 * u0 * (self.nx / (1.42 - (-0.42)))
        out = dict(fx=fx, fy=fy, surf=int(surf), rev=int(rev), any=surf > 0)
        if surf > 0:
            cl, cd = fy / q, fx / q
            self.cl_smooth = cl if self.cl_smooth is None else self.cl_smooth * 0.9 + cl * 0.1
            self.cd_smooth = cd if self.cd_smooth is None else self.cd_smooth * 0.9 + cd * 0.1
            self.sep_frac = self.sep_frac * 0.85 + (rev / surf) * 0.15
            out.update(CL_raw=cl, CD_raw=cd)
        out.update(CL=self.cl_smooth, CD=self.cd_smooth, sep_frac=self.sep_frac)
        if self.t.steps > 0:
            me = self.comm.allreduce(self.t.me_history(1)[0], "sum")
            fxm, fym = float(me[0]) / 2.0 ** 40, float(me[1]) / 2.0 ** 40
            out.update(Fx_me=fxm, Fy_me=fym, CL_me=fym / q, CD_me=fxm / q)
        return out

    def state_hash(self) -> np.ndarray:
        """Nine uint64 checksum words of the whole lattice's populations: the slabs' words added
        modulo 2^64 (identical for every decomposition of the same state)."""
        mine = self.t.state_hash()
        # int64 two's-complement addition wraps exactly like uint64 addition
        tot = self.comm.allreduce(mine.view(np.int64), "sum")
        return np.ascontiguousarray(tot, dtype=np.int64).view(np.uint64)

    def _sticky(self) -> dict:
        return dict(maxS=self.max_s, cpMin=self.cp_min, cpMax=self.cp_max,
                    cl_smooth=self.cl_smooth if self.cl_smooth is not None else 0.0,
                    cd_smooth=self.cd_smooth if self.cd_smooth is not None else 0.0,
                    sep_frac=self.sep_frac, ema_valid=self.cl_smooth is not None)

    def run_frames(self, nframes: int, controls=None, steps_per_frame: int = 4, forces_every: int = 3) -> dict:
        """``frame()`` x nframes (HTML:902-930) on the decomposed lattice with no host synchronisation
        and no collective inside the loop: every slab enqueues all its frames (``alb_frames_enqueue``),
        the per-frame partial reductions land in host memory as the frames complete, and only then
        are they gathered (ONE all-gather) and run through the page's host logic
        (:func:`combine_frame_partials`).  Same columns as ``WindTunnel.run_frames``."""
        if self.halo == "nccl":
            raise NotImplementedError("run_frames needs the in-kernel halo (halo='p2p'); use frame() with halo='nccl'")
        self.t.frames_enqueue(nframes, controls=controls, steps_per_frame=steps_per_frame, forces_every=forces_every)
        mine = self.t.frames_collect_raw()
        parts = self.comm.allgather_array(mine)
        if self.comm.world == 1 and self.t.ny_local == self.ny:
            return {k: mine[:, i] for i, k in enumerate(FRAME_COLUMNS)}       # whole lattice: finished records
        st = self._sticky()
        series = combine_frame_partials(parts, st)
        self.max_s, self.cp_min, self.cp_max = st["maxS"], st["cpMin"], st["cpMax"]
        if st["ema_valid"]:
            self.cl_smooth, self.cd_smooth = st["cl_smooth"], st["cd_smooth"]
        self.sep_frac = st["sep_frac"]
        self.t.set_stats(self.max_s, self.cp_min, self.cp_max)
        return series

    # -- field modes (HTML:395-422, 527-545) on the decomposed lattice ---------------------------
    def _exchange_macro_edges(self):
        """The vorticity taps of a slab's first / last row reach into the neighbouring slab: every
        rank publishes ux, uy of its two edge rows (2 x 2 x nx floats) and installs its neighbours'."""
        lo, hi = self.t.macro_edges()
        edges = self.comm.allgather_array(np.stack([lo, hi]).astype(np.float64))     # (world, 2, 2, nx)
        r, w = self.comm.rank, self.comm.world
        below = edges[r - 1, 1].astype(np.float32) if r > 0 else None
        above = edges[r + 1, 0].astype(np.float32) if r < w - 1 else None
        self.t.set_macro_ghosts(below, above)

    def field(self, mode="speed") -> np.ndarray:
        """This rank's rows of the render shader's scalar (NaN in solids); uses the lattice-wide sticky
        autoscale values of the last ``update_stats`` / ``run_frames``."""
        from .tunnel import FIELD_MODES
        if FIELD_MODES[mode] == 2 and self.comm.world > 1:
            self._exchange_macro_edges()
        return self.t.field(mode)

    def rgba(self, mode="speed") -> np.ndarray:
        """This rank's rows of the colour-mapped field (page palettes, HTML:371-393), (ny_local, nx, 4)."""
        from .tunnel import FIELD_MODES
        if FIELD_MODES[mode] == 2 and self.comm.world > 1:
            self._exchange_macro_edges()
        return self.t.rgba(mode)

    def stall_state(self) -> str:
        """Text of the separation card (HTML:869-884) from the lattice-wide separation fraction."""
        x = self.sep_frac * 100
        pct = math.floor(x) + (1 if x - math.floor(x) >= 0.5 else 0)      # Math.round, HTML:869
        return "Attached" if pct < 5 else (f"{pct}% sep" if pct < 25 else f"STALL \u2248 {pct}% sep")

    def reynolds(self) -> float:
        return self.t.reynolds()

    def frame(self) -> dict:
        """The reference frame (HTML:902-930) across slabs: 4 steps, autoscale, forces every 3rd."""
        self.step(4)
        out = {"stats": self.update_stats()}
        self._frames = getattr(self, "_frames", 0) + 1
        if self._frames % 3 == 0:
            out["forces"] = self.forces()
        return out

    def gather(self, what: str = "macro"):
        """Assemble whole-lattice arrays on rank 0 (tests / dumps)."""
        if what == "macro":
            parts = [self.comm.gather_arrays(a) for a in self.t.macro()]
            return None if parts[0] is None else tuple(np.concatenate(p, axis=0) for p in parts)
        if what == "populations":
            p = self.comm.gather_arrays(self.t.populations())
            return None if p is None else np.concatenate(p, axis=1)
        if what.startswith("field:") or what.startswith("rgba:"):
            kind, mode = what.split(":")
            p = self.comm.gather_arrays(self.field(mode) if kind == "field" else self.rgba(mode))
            return None if p is None else np.concatenate(p, axis=0)
        if what == "mask":
            p = self.comm.gather_arrays(self.t.mask())
            return None if p is None else np.concatenate(p, axis=0)
        raise ValueError(what)


def bench_e2e(tun: DistributedTunnel, comm: Comm, steps: int, cells_global: int) -> dict:
    """End-to-end rate through the public API with host-side control, as a user drives it.

    The reference's frame loop (HTML:902-930): every frame the host supplies the control inputs
    (U0, tau -- the sliders), 4 steps run, the autoscale statistics are refreshed, every 3rd frame
    the pressure forces and their EMAs are updated, and the frame's 12-double record lands in host
    memory.  One GPU: ``WindTunnel.run_frames`` (C ABI ``alb_run_frames``; the loop is enqueued
    asynchronously, records are copied device->host as frames complete, one synchronisation at
    the end).  Several GPUs: the same frame driven from Python with cross-rank reductions of the
    partial sums every frame.  Host wall clock around the call(s), max over ranks.
    """
    nframes = max(3, steps // 4)
    u0, tau = tun.t.params()
    if comm.world == 1:
        controls = np.tile(np.array([u0, tau]), (nframes, 1))
        tun.t.run_frames(3, controls=controls[:3])
        tun.sync()
        t0 = time.perf_counter()
        series = tun.t.run_frames(nframes, controls=controls)      # synchronises at the end
        dt = time.perf_counter() - t0
        assert series["CL"].shape == (nframes,)
        h2d, d2h = 16.0, 12 * 8.0
        what = ("alb_run_frames: per frame 16 B of control inputs (U0, tau; kernel arguments), 4 steps, on-device "
                "statistics/force EMAs, a 96 B record copied to host memory; no host synchronisation inside the "
                "loop; host wall clock around the call including the final synchronisation")
    else:
        controls = np.tile(np.array([u0, tau]), (nframes, 1))
        if tun.halo == "nccl":
            for _ in range(3):
                tun.frame()
            tun.sync()
            comm.barrier()
            t0 = time.perf_counter()
            for _ in range(nframes):
                tun.set_params(u0, tau)
                tun.frame()
            tun.sync()
            dt = comm.max_float(time.perf_counter() - t0)
            comm.barrier()
            h2d, d2h = 16.0, 3 * 8.0 + (4 * 8.0 + 16.0) / 3.0
            what = ("frame loop driven from Python on every rank: set_params + 4 steps + per-slab statistics read "
                    "back and all-reduced every frame, forces every 3rd frame; host wall clock, max over ranks")
        else:
            tun.run_frames(3, controls=controls[:3])
            tun.sync()
            comm.barrier()
            t0 = time.perf_counter()
            series = tun.run_frames(nframes, controls=controls)
            dt = comm.max_float(time.perf_counter() - t0)
            comm.barrier()
            assert series["CL"].shape == (nframes,)
            h2d, d2h = 16.0, 12 * 8.0
            what = ("DistributedTunnel.run_frames: every rank enqueues all frames on its slab (alb_frames_enqueue: per "
                    "frame 16 B of control inputs, 4 steps with in-kernel NVLink halo, on-device partial reductions, a "
                    "96 B record copied to host memory), no host synchronisation or collective inside the loop; then one "
                    "all-gather of the records and the page's EMA logic on the host; host wall clock around the call "
                    "including the final synchronisation and the all-gather, max over ranks")
    return {"value": cells_global * 4 * nframes / dt / 1e9, "unit": "GLUPS",
            "h2d_bytes_per_step": h2d / 4.0, "d2h_bytes_per_step": d2h / 4.0,
            "frames": nframes, "steps_per_frame": 4, "what": what}


def bench_e2e_fields(tun: DistributedTunnel, cells_global: int, nframes: int = 3) -> dict:
    """The frame loop WITH the frame's image, as the page draws it every frame (HTML:909
    ``renderField``): per frame 4 steps, the autoscale statistics, the colour-mapped speed field
    rendered on the device (``alb_get_rgba``) and copied to host memory.  One GPU.  At configs[3] the
    image is 2.1 GB per frame, so this figure is a PCIe number; the headless loop is ``e2e``."""
    t = tun.t
    u0, tau = t.params()
    t.frame(want_field=None)
    t.rgba("speed")
    t.sync()
    t0 = time.perf_counter()
    for _ in range(nframes):
        t.set_params(u0, tau)
        t.run_frames(1)
        img = t.rgba("speed")
    dt = time.perf_counter() - t0
    assert img.shape == (t.ny_local, t.nx, 4)
    return {"value": cells_global * 4 * nframes / dt / 1e9, "unit": "GLUPS", "frames": nframes, "steps_per_frame": 4,
            "h2d_bytes_per_step": 16.0 / 4.0, "d2h_bytes_per_step": (12 * 8.0 + 4.0 * cells_global) / 4.0,
            "what": "per frame: 4 steps + on-device statistics (alb_run_frames), then alb_get_rgba: macroscopic pass, "
                    "render kernel with the page's palette, RGBA8 image copied to pageable host memory; host wall clock"}

"""CPU restatement of the tunnel's tracer particles (ORACLE, tests only).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  ``sampleScalar/sampleUV/advect`` are pinned against the
reference's own JavaScript executed by tests/refexec (tests/test_reference_pins.py); spawning uses
``Math.random`` in the reference, so there only the algorithm can be mirrored, not a run.

Follows pages/airfoil_flow_lbm_aerolab.html: ``sampleScalar``/``sampleUV`` 616-639, ``spawn``
730-736, ``initParts`` 737-753, ``advect`` 754-767, ``stepParticles`` 780-808 (without the canvas
strokes), the trail-count slider 961-967.  ``Math.random()`` is replaced by the same
counter-based generator the CUDA library uses (splitmix64 of seed, particle index and draw
counter) so that the two can be compared particle by particle.
"""
from __future__ import annotations

import math

import numpy as np

from .geometry import DX0, DX1, DY0, DY1

M64 = (1 << 64) - 1


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


class Particles:
    STALL_SPEED2 = 3e-6      # HTML:777
    STALL_DRAIN = 0.18       # HTML:778

    def __init__(self, n: int, seed: int):
        self.seed = seed
        self.p = []
        self.ctr = []
        for i in range(n):                                   # initParts(), HTML:737-753
            self.ctr.append(0)
            if self._rand(i) < 0.35:
                c, half = (DY0 + DY1) / 2, (DY1 - DY0) / 6
                lane = c + (self._rand(i) - 0.5) * 2 * half
            else:
                lane = DY0 + ((i + 0.5) / n) * (DY1 - DY0) + (self._rand(i) - 0.5) * 0.003
            q = self._spawn(i, True, lane)
            q["life"] *= self._rand(i)
            q["x"] = DX0 + self._rand(i) * (DX1 - DX0) * 0.95
            self.p.append(q)

    def _rand(self, pid: int) -> float:
        k = splitmix64(self.seed ^ splitmix64(((pid & 0xffffffff) << 32) | self.ctr[pid]))
        self.ctr[pid] += 1
        return (k >> 11) * 2.0 ** -53

    def _spawn(self, pid, edge, lane=None):                  # spawn(), HTML:730-736
        if lane is None:
            lane = DY0 + self._rand(pid) * (DY1 - DY0)
        if edge or self._rand(pid) < 0.82:
            return dict(x=DX0 + 0.001, y=lane, life=220 + self._rand(pid) * 300, lane=lane)
        return dict(x=DX0 + self._rand(pid) * (DX1 - DX0), y=DY0 + self._rand(pid) * (DY1 - DY0),
                    life=150 + self._rand(pid) * 250, lane=lane)

    def resize(self, n: int):                                # HTML:961-967
        while len(self.p) > n:
            self.p.pop()
            self.ctr.pop()
        while len(self.p) < n:
            self.ctr.append(0)
            self.p.append(self._spawn(len(self.p), False))

    @staticmethod
    def _sample(field, mask, wx, wy):                        # sampleScalar(), HTML:616-632
        ny, nx = mask.shape
        if wx < DX0 or wx > DX1 or wy < DY0 or wy > DY1:
            return None
        fx = (wx - DX0) / (DX1 - DX0) * nx - 0.5
        fy = (wy - DY0) / (DY1 - DY0) * ny - 0.5
        ix = max(0, min(math.floor(fx), nx - 2))
        iy = max(0, min(math.floor(fy), ny - 2))
        tx, ty = fx - ix, fy - iy
        ws = [(1 - tx) * (1 - ty), tx * (1 - ty), (1 - tx) * ty, tx * ty]
        cs = [(iy, ix), (iy, ix + 1), (iy + 1, ix), (iy + 1, ix + 1)]
        s = w = 0.0
        for k in range(4):
            if not mask[cs[k]] and math.isfinite(field[cs[k]]):
                s += float(field[cs[k]]) * ws[k]
                w += ws[k]
        return s / w if w > 0 else None

    @classmethod
    def advect(cls, U, V, mask, x, y, dt):
        """advect(), HTML:754-767: (nx, ny, speed) or None."""
        u1 = cls._sample(U, mask, x, y)
        v1 = cls._sample(V, mask, x, y)
        if u1 is None or v1 is None:
            return None
        k_base = 0.00105 * dt
        speed1 = math.hypot(u1, v1)
        dt_eff = k_base
        max_disp = 0.05
        if speed1 * dt_eff > max_disp:
            dt_eff = max_disp / max(speed1, 1e-6)
        midx, midy = x + u1 * dt_eff * 0.5, y + v1 * dt_eff * 0.5
        u2 = cls._sample(U, mask, midx, midy)
        v2 = cls._sample(V, mask, midx, midy)
        if u2 is None or v2 is None:
            u2, v2 = u1, v1
        return (x + u2 * dt_eff, y + v2 * dt_eff, math.hypot(u2, v2))

    def step(self, dt, mask, ux, uy, u0):
        """stepParticles(dt), HTML:780-808; returns (n, 8) like alb_particles_get."""
        with np.errstate(all="ignore"):
            U = (ux.astype(np.float64) / u0).astype(np.float32)      # Ufield, HTML:603-604
            V = (uy.astype(np.float64) / u0).astype(np.float32)
        out = np.zeros((len(self.p), 8))
        for i, p in enumerate(self.p):
            adv = self.advect(U, V, mask, p["x"], p["y"], dt)
            stalled = adv is not None and adv[2] * adv[2] < self.STALL_SPEED2
            p["life"] -= dt * (self.STALL_DRAIN if stalled else 0.06)
            if adv is None or p["life"] <= 0:
                q = self._spawn(i, True, p["lane"])
                self.p[i] = q
                out[i] = (q["x"], q["y"], q["life"], q["lane"], q["x"], q["y"], 0.0, 1.0)
                continue
            x0, y0 = p["x"], p["y"]
            p["x"], p["y"] = adv[0], adv[1]
            out[i] = (p["x"], p["y"], p["life"], p["lane"], x0, y0, adv[2], 0.0)
        return out

    def table(self):
        return np.array([[q["x"], q["y"], q["life"], q["lane"]] for q in self.p])

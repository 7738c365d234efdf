"""Second, independently written restatement of the reference step (NumPy, fp32).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Pinned through oracle/lbm_ref.c (see oracle/__init__.py).


Purpose: guard against a reading error in ``oracle/lbm_ref.c``.  The two are
written differently (scalar C with branches versus whole-array NumPy with
``np.where``) and ``tests/test_oracle.py`` requires them to agree bit for bit.
NumPy evaluates every float32 operation as a separate correctly rounded IEEE
operation, so no FMA can sneak in.

Follows ``STEP_FS_SRC`` HTML:222-360 and ``equilibriumInitData`` HTML:474-490.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
E = [(0, 0), (1, 0), (0, 1), (-1, 0), (0, -1), (1, 1), (-1, 1), (-1, -1), (1, -1)]
OPP = [0, 3, 4, 1, 2, 7, 8, 5, 6]
W0 = f32(4.0) / f32(9.0)
WS = f32(1.0) / f32(9.0)
WD = f32(1.0) / f32(36.0)
W = [W0, WS, WS, WS, WS, WD, WD, WD, WD]


def feq(i, rho, ux, uy):
    ex, ey = E[i]
    eu = f32(ex) * ux + f32(ey) * uy
    uu = ux * ux + uy * uy
    return W[i] * rho * (f32(1.0) + f32(3.0) * eu + f32(4.5) * eu * eu - f32(1.5) * uu)


def init(nx, ny, u0):
    u0 = float(u0)
    w = [4 / 9, 1 / 9, 1 / 9, 1 / 9, 1 / 9, 1 / 36, 1 / 36, 1 / 36, 1 / 36]
    F = np.empty((9, ny, nx), f32)
    for i, (ex, _) in enumerate(E):
        eu = ex * u0
        uu = u0 * u0
        F[i] = f32(w[i] * (1 + 3 * eu + 4.5 * eu * eu - 1.5 * uu))
    return F


def _shift(a, ex, ey):
    """a sampled at (x-ex, y-ey); edges wrap but are never used by interior cells."""
    return np.roll(a, shift=(ey, ex), axis=(0, 1))


def step(mask, F, tau, u0):
    """Returns (Fnew, rho, ux, uy)."""
    tau = f32(tau)
    u0 = f32(u0)
    _, ny, nx = F.shape
    solid = mask > 0
    X = np.arange(nx)[None, :].repeat(ny, 0)
    Y = np.arange(ny)[:, None].repeat(nx, 1)
    outlet = (~solid) & (X == nx - 1)
    equil = (~solid) & (~outlet) & ((X == 0) | (Y == ny - 1) | (Y == 0))

    with np.errstate(all="ignore"):
        # interior everywhere (garbage where another branch wins)
        fin = []
        for i, (ex, ey) in enumerate(E):
            src_solid = _shift(solid, ex, ey)
            fin.append(np.where(src_solid, F[OPP[i]], _shift(F[i], ex, ey)))
        rho = np.zeros((ny, nx), f32)
        for i in range(9):
            rho = rho + fin[i]
        ux = (fin[1] + fin[5] + fin[8] - fin[3] - fin[6] - fin[7]) / rho
        uy = (fin[2] + fin[5] + fin[6] - fin[4] - fin[7] - fin[8]) / rho
        rho = np.minimum(np.maximum(rho, f32(0.5)), f32(2.0))
        spd2 = ux * ux + uy * uy
        umax = f32(0.35)
        fast = spd2 > umax * umax
        k = umax / np.sqrt(np.where(fast, spd2, f32(1.0)))
        ux = np.where(fast, ux * k, ux)
        uy = np.where(fast, uy * k, uy)
        out = np.empty_like(F)
        for i in range(9):
            eq = feq(i, rho, ux, uy)
            out[i] = fin[i] - (fin[i] - eq) / tau

        # outlet: copy from x-1 of the source state
        Fm = np.roll(F, 1, axis=2)
        rho_o = Fm[0] + Fm[1] + Fm[2] + Fm[3] + Fm[4] + Fm[5] + Fm[6] + Fm[7] + Fm[8]
        ux_o = (Fm[1] + Fm[5] + Fm[8] - Fm[3] - Fm[6] - Fm[7]) / rho_o
        uy_o = (Fm[2] + Fm[5] + Fm[6] - Fm[4] - Fm[7] - Fm[8]) / rho_o

    one = f32(1.0)
    zero = f32(0.0)
    for i in range(9):
        out[i] = np.where(outlet, Fm[i], out[i])
        out[i] = np.where(equil, feq(i, one, u0, zero), out[i])
        out[i] = np.where(solid, F[OPP[i]], out[i])
    rho = np.where(outlet, rho_o, rho)
    ux = np.where(outlet, ux_o, ux)
    uy = np.where(outlet, uy_o, uy)
    rho = np.where(equil, one, rho)
    ux = np.where(equil, u0, ux)
    uy = np.where(equil, zero, uy)
    rho = np.where(solid, one, rho)
    ux = np.where(solid, zero, ux)
    uy = np.where(solid, zero, uy)
    return out, rho.astype(f32), ux.astype(f32), uy.astype(f32)

"""Host-side airfoil shapes of the tunnel's control surface (float64).

Product code (not the oracle): the built-in shape generators the reference page
offers next to injected user coordinates, and the 6-decimal rounding its Python
bridge applies.  Rotation, panelisation and scan conversion live in the CUDA
library (``alb_rasterize``); only the raw chord-normalised outline is produced
here.

Reference: pages/airfoil_flow_lbm_aerolab.html:99-129 (``naca4``, ``clarkY``,
``SHAPES``) and pages/Airfoil_Analysis.py:34-36 (rounding).
"""
from __future__ import annotations

from math import atan, cos, pi, sin, sqrt

# world window of the tunnel, HTML:73
DX0, DX1, DY0, DY1 = -0.42, 1.42, -0.46, 0.46


def naca4(m: float, p: float, t: float, n: int = 50):
    """NACA 4-digit outline, TE -> upper -> LE -> lower -> TE (2n+1 points).

    Cosine spacing in x, thickness applied normal to the camber line, closed
    trailing edge (last thickness coefficient -0.1036).  HTML:99-116.
    """
    m, p, t = m / 100, p / 10, t / 100
    upper, lower = [], []
    for i in range(n + 1):
        beta = pi * i / n
        x = 0.5 * (1 - cos(beta))
        yt = 5 * t * (0.2969 * sqrt(x) - 0.126 * x - 0.3516 * x * x + 0.2843 * x ** 3 - 0.1036 * x ** 4)
        yc = dyc = 0.0
        if m > 0:
            if x < p:
                yc = m / p / p * (2 * p * x - x * x)
                dyc = 2 * m / p / p * (p - x)
            else:
                yc = m / (1 - p) ** 2 * ((1 - 2 * p) + 2 * p * x - x * x)
                dyc = 2 * m / (1 - p) ** 2 * (p - x)
        th = atan(dyc)
        upper.append([x - yt * sin(th), yc + yt * cos(th)])
        lower.append([x + yt * sin(th), yc - yt * cos(th)])
    return upper[::-1] + lower[1:]


_CLARK_Y_PERCENT = (
    (100, .44), (95, 1.46), (90, 2.22), (80, 3.69), (70, 5.07), (60, 6.23), (50, 7.1), (40, 7.62),
    (30, 7.79), (25, 7.67), (20, 7.35), (15, 6.79), (10, 5.88), (7.5, 5.23), (5, 4.39), (2.5, 3.18),
    (1.25, 2.17), (0, 0), (1.25, -1.35), (2.5, -1.93), (5, -2.55), (7.5, -2.9), (10, -3.05),
    (15, -3.01), (20, -2.75), (25, -2.41), (30, -2.06), (40, -1.38), (50, -.85), (60, -.44),
    (70, -.16), (80, 0), (90, 0), (95, 0), (100, -.44))


def clark_y():
    """Clark-Y table (percent chord / 100), open trailing edge.  HTML:118-121."""
    return [[x / 100, y / 100] for x, y in _CLARK_Y_PERCENT]


def naca_digits(code: str, n: int = 50):
    """``"2412"`` -> naca4(2, 4, 12, n)."""
    code = code.strip().lower().replace("naca", "").strip()
    if len(code) != 4 or not code.isdigit():
        raise ValueError(f"not a NACA 4-digit designation: {code!r}")
    return naca4(int(code[0]), int(code[1]), int(code[2:]), n)


# HTML:123-129
SHAPES = {
    "naca0012": lambda: naca4(0, 0, 12, 50),
    "naca2412": lambda: naca4(2, 4, 12, 50),
    "naca4412": lambda: naca4(4, 4, 12, 50),
    "naca6409": lambda: naca4(6, 4, 9, 50),
    "clark_y": clark_y,
}


def round_coords(coords):
    """What ``build_lbm_component`` injects: 6-decimal rounding (AA.py:34-36)."""
    return [[round(float(x), 6), round(float(y), 6)] for x, y in coords]

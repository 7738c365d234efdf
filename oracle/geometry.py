"""Float64 restatement of the reference tunnel's geometry pipeline (ORACLE).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Pinned against the reference's own JavaScript executed by tests/refexec (tests/test_reference_pins.py):
panel nodes bitwise, masks identical.

Follows ``pages/airfoil_flow_lbm_aerolab.html`` (cited as HTML:line):

* ``naca4``           HTML:99-116
* ``clark_y``         HTML:118-121
* ``SHAPES``          HTML:123-129
* ``rotate``          HTML:133-140
* ``panelise``        HTML:142-157
* ``raster_mask``     HTML:160-182
* ``build_geometry``  HTML:559-577 (selection between injected coordinates
  and the built-in shapes, rotate -> panelise -> rasterise)
* ``round_coords``    pages/Airfoil_Analysis.py:34-36 (6-decimal rounding on
  injection)

Everything is scalar Python float (IEEE binary64) arithmetic written in the
reference's operand order; JavaScript ``Math.*`` maps to Python ``math.*``
(glibc libm), which is the stated definition of "bit-exact" for the mask.
"""
from __future__ import annotations

import math

import numpy as np

# HTML:73
DX0, DX1, DY0, DY1 = -0.42, 1.42, -0.46, 0.46
# HTML:131
NP = 160


def naca4(m, p, t, n):
    """HTML:99-116. Closed loop TE -> upper -> LE -> lower -> TE, 2n+1 points."""
    m = m / 100
    p = p / 10
    t = t / 100
    up = []
    lo = []
    for i in range(n + 1):
        b = math.pi * i / n
        x = 0.5 * (1 - math.cos(b))
        yt = 5 * t * (0.2969 * math.sqrt(x) - 0.126 * x - 0.3516 * x * x
                      + 0.2843 * x ** 3 - 0.1036 * x ** 4)
        yc = 0.0
        dyc = 0.0
        if m > 0:
            if x < p:
                yc = m / p / p * (2 * p * x - x * x)
                dyc = 2 * m / p / p * (p - x)
            else:
                yc = m / (1 - p) ** 2 * ((1 - 2 * p) + 2 * p * x - x * x)
                dyc = 2 * m / (1 - p) ** 2 * (p - x)
        th = math.atan(dyc)
        up.append([x - yt * math.sin(th), yc + yt * math.cos(th)])
        lo.append([x + yt * math.sin(th), yc - yt * math.cos(th)])
    up.reverse()
    return up + lo[1:]


_CLARK_Y = [[100, .44], [95, 1.46], [90, 2.22], [80, 3.69], [70, 5.07], [60, 6.23],
            [50, 7.1], [40, 7.62], [30, 7.79], [25, 7.67], [20, 7.35], [15, 6.79],
            [10, 5.88], [7.5, 5.23], [5, 4.39], [2.5, 3.18], [1.25, 2.17], [0, 0],
            [1.25, -1.35], [2.5, -1.93], [5, -2.55], [7.5, -2.9], [10, -3.05],
            [15, -3.01], [20, -2.75], [25, -2.41], [30, -2.06], [40, -1.38],
            [50, -.85], [60, -.44], [70, -.16], [80, 0], [90, 0], [95, 0], [100, -.44]]


def clark_y():
    """HTML:118-121. 35-point table / 100, open trailing edge."""
    return [[x / 100, y / 100] for x, y in _CLARK_Y]


# HTML:123-129
SHAPES = {
    "naca0012": lambda: naca4(0, 0, 12, 50),
    "naca2412": lambda: naca4(2, 4, 12, 50),
    "naca4412": lambda: naca4(4, 4, 12, 50),
    "naca6409": lambda: naca4(6, 4, 9, 50),
    "clark_y": clark_y,
}


def round_coords(coords):
    """pages/Airfoil_Analysis.py:34-36: what the Python bridge injects."""
    return [[round(float(x), 6), round(float(y), 6)] for x, y in coords]


def rotate(coords, a_deg):
    """HTML:133-140. Rotation by -a_deg about (0.25, 0)."""
    a = -a_deg * math.pi / 180
    ca = math.cos(a)
    sa = math.sin(a)
    px = 0.25
    py = 0
    out = []
    for x, y in coords:
        dx = x - px
        dy = y - py
        out.append([px + dx * ca - dy * sa, py + dx * sa + dy * ca])
    return out


def panelise(coords):
    """HTML:142-157. NP+1 cosine-spaced arc-length samples of the polyline."""
    xs = [p[0] for p in coords]
    ys = [p[1] for p in coords]
    arc = [0.0]
    for i in range(1, len(coords)):
        arc.append(arc[i - 1] + math.hypot(xs[i] - xs[i - 1], ys[i] - ys[i - 1]))
    L = arc[-1]
    xp = []
    yp = []
    for i in range(NP + 1):
        s = L * 0.5 * (1 - math.cos(math.pi * i / NP))
        j = 0
        while j < len(arc) - 2 and arc[j + 1] < s:
            j += 1
        t = (s - arc[j]) / (arc[j + 1] - arc[j] + 1e-12)
        xp.append(xs[j] + (xs[j + 1] - xs[j]) * t)
        yp.append(ys[j] + (ys[j + 1] - ys[j]) * t)
    return xp, yp


def raster_rows(xp, yp, nx, ny, rows):
    """HTML:160-182 for the lattice rows listed in ``rows`` (global indices; rows outside
    [0, ny) come out empty): uint8[len(rows), nx] of 0/255.  Scanline even-odd fill.

    Row 0 is the bottom of the world window (y up).  y is sampled at cell
    centres, x at integer node positions; the polygon is NOT closed and an
    unpaired last crossing is dropped -- all as in the reference.
    """
    rows = list(rows)
    mask = np.zeros((len(rows), nx), dtype=np.uint8)
    n = len(xp)
    for r, iy in enumerate(rows):
        if iy < 0 or iy >= ny:
            continue
        wy = DY0 + (iy + 0.5) / ny * (DY1 - DY0)
        xs = []
        for i in range(n - 1):
            y1 = yp[i]
            y2 = yp[i + 1]
            if (y1 > wy) != (y2 > wy):
                x1 = xp[i]
                x2 = xp[i + 1]
                xs.append(x1 + (x2 - x1) * (wy - y1) / (y2 - y1))
        xs.sort()
        k = 0
        while k + 1 < len(xs):
            ix0 = math.ceil((xs[k] - DX0) / (DX1 - DX0) * nx)
            ix1 = math.floor((xs[k + 1] - DX0) / (DX1 - DX0) * nx)
            ix0 = max(0, ix0)
            ix1 = min(nx - 1, ix1)
            if ix1 >= ix0:
                mask[r, ix0:ix1 + 1] = 255
            k += 2
    return mask


def raster_mask(xp, yp, nx, ny):
    """HTML:160-182 for the whole lattice: uint8[ny, nx] of 0/255."""
    return raster_rows(xp, yp, nx, ny, range(ny))


def build_geometry(base_coords, a_deg, nx, ny):
    """HTML:559-577 restricted to what the LBM uses: rotate -> panelise -> mask."""
    coords = rotate(base_coords, a_deg)
    xp, yp = panelise(coords)
    return xp, yp, raster_mask(xp, yp, nx, ny)


def shape_coords(shape_key=None, user_coords=None):
    """HTML:561-563: injected coordinates win over the built-in shapes."""
    if user_coords is not None and len(user_coords) > 0:
        return [list(map(float, p)) for p in user_coords]
    return SHAPES[shape_key]()

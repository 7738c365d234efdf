"""Regenerate the committed golden fixtures (run in the BUILD container only).

    python tests/golden/make_golden.py

Two sources:
  * the CPU oracle (oracle/, a restatement -- the reference's LBM cannot run
    here, PARITY UNPINNED): mask hashes, the configs[0] field dump after 1,000
    steps, force values;
  * the REAL reference parser, imported from /root/reference/main.py (with a
    stand-in for the missing `slowapi` package): parsed coordinates of the
    reference's own test fixture (test_main.py:33-50) and of a Lednicer file.
    /root/reference does not exist on the GPU box, hence the committed JSON.
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "airfoil-cfd-tool_b200"))

from oracle import geometry as ogeo  # noqa: E402
from oracle import lbm as olbm  # noqa: E402

MASK_CASES = [
    ("naca0012", 5.0, 320, 160), ("naca2412", 6.0, 320, 160), ("naca4412", 10.0, 320, 160),
    ("naca6409", -7.5, 320, 160), ("clark_y", 6.0, 320, 160), ("clark_y", 6.0, 2048, 1024),
    ("naca0012", 0.0, 2048, 1024), ("naca0012", 20.0, 2048, 1024), ("naca4412", 10.0, 4096, 2048),
    ("naca2412", 5.0, 333, 171), ("naca2412", 5.0, 32768, 16384),
]

NACA0012_SELIG = """NACA 0012
1.000000  0.001260
0.933013  0.005740
0.750000  0.015970
0.500000  0.030230
0.250000  0.041210
0.066987  0.031530
0.000000  0.000000
0.066987 -0.031530
0.250000 -0.041210
0.500000 -0.030230
0.750000 -0.015970
0.933013 -0.005740
1.000000 -0.001260
"""

LEDNICER = """NACA 2412 (Lednicer)
  11.   11.
 0.000000  0.000000
 0.050000  0.034000
 0.100000  0.047000
 0.200000  0.063000
 0.300000  0.071000
 0.400000  0.073000
 0.500000  0.070000
 0.600000  0.062000
 0.700000  0.050000
 0.850000  0.028000
 1.000000  0.001300

 0.000000  0.000000
 0.050000 -0.021000
 0.100000 -0.027000
 0.200000 -0.033000
 0.300000 -0.035000
 0.400000 -0.034000
 0.500000 -0.031000
 0.600000 -0.026000
 0.700000 -0.020000
 0.850000 -0.010000
 1.000000 -0.001300
"""


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def masks():
    out = []
    for shape, alpha, nx, ny in MASK_CASES:
        _, _, m = ogeo.build_geometry(ogeo.SHAPES[shape](), alpha, nx, ny)
        out.append(dict(shape=shape, alpha=alpha, nx=nx, ny=ny, solid=int((m > 0).sum()), sha256=sha(m)))
        print(out[-1])
    return out


def default_case():
    o = olbm.OracleTunnel(320, 160)
    o.apply_geometry(ogeo.SHAPES["naca0012"](), 5.0)
    hashes = {}
    for s in range(1, 1001):
        o.step(1)
        if s in (1, 2, 10, 100, 1000):
            hashes[str(s)] = dict(f=sha(o.F), rho=sha(o.rho), ux=sha(o.ux), uy=sha(o.uy))
    f = o.compute_forces()
    cl_me, cd_me = o.me_coeffs()
    meta = dict(case="NACA 0012 alpha=5 320x160 U0=0.06 tau=0.58", hashes=hashes,
                mass=olbm.total_mass(o.F), CL_raw=float(f["CL_raw"]), CD_raw=float(f["CD_raw"]),
                surf=f["surf"], rev=f["rev"], CL_me=cl_me, CD_me=cd_me, clamp_hits=o.clamp_hits,
                me_first8=[list(map(int, v)) for v in o.me_hist[:8]],
                me_last=list(map(int, o.me_hist[-1])))
    np.savez_compressed(os.path.join(HERE, "config0_step1000.npz"), rho=o.rho, ux=o.ux, uy=o.uy,
                        mask=o.mask)
    return meta


def reference_parser():
    from aerolab_lbm.dat import load_reference_parser
    parse = load_reference_parser("/root/reference")
    out = {}
    for name, text in (("naca0012_selig_test_main", NACA0012_SELIG), ("naca2412_lednicer", LEDNICER)):
        with tempfile.NamedTemporaryFile("w", suffix=".dat", delete=False) as fh:
            fh.write(text)
        coords, fixes = parse(fh.name)
        os.unlink(fh.name)
        out[name] = dict(text=text, coords=[[float(x), float(y)] for x, y in coords], fixes=list(fixes))
        print(name, len(coords), fixes)
    return out


if __name__ == "__main__":
    data = dict(masks=masks(), config0=default_case(), parser=reference_parser())
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(data, fh, indent=1)
    print("wrote golden.json and config0_step1000.npz")

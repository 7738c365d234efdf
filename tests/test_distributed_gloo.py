"""CPU, world_size 2, gloo: the N>1 host path (slab partition, halo exchange, reductions)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, assert_bitwise
from oracle import geometry as ogeo
from oracle import lbm as olbm


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,ny", [(2, 64), (3, 50)])
def test_gloo_slabs_match_whole_lattice(tmp_path, world, ny):
    nx, nsteps = 96, 25
    out = str(tmp_path / "result.npz")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()), WORLD_SIZE=str(world),
               OMP_NUM_THREADS="2")
    nframes = 7
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_worker.py"), str(nx), str(ny),
                               str(nsteps), out, str(nframes)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(world)]
    logs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        logs.append(o.decode(errors="replace"))
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    o = olbm.OracleTunnel(nx, ny)
    o.apply_geometry(ogeo.SHAPES["naca4412"](), 10.0)
    o.step(nsteps)
    # the frame loop of the decomposed lattice (per-slab partial records -> all-gather ->
    # combine_frame_partials) against the page's frame loop on the whole lattice
    rows = []
    for f in range(1, nframes + 1):
        o.step(4)
        o.update_fields()
        row = dict(maxS=o.max_s, cpMin=o.cp_min, cpMax=o.cp_max, CL=np.nan, CD=np.nan, sep=o.sep_frac, CL_me=o.me_coeffs()[0])
        if f % 3 == 0:
            o.compute_forces()
        if o.cl_smooth is not None:
            row.update(CL=o.cl_smooth, CD=o.cd_smooth)
        row["sep"] = o.sep_frac
        rows.append(row)
    r = np.load(out)
    for k, row in enumerate(rows):
        assert (r["series_maxS"][k], r["series_cpMin"][k], r["series_cpMax"][k]) == (row["maxS"], row["cpMin"], row["cpMax"]), k
        assert r["series_CL_me"][k] == pytest.approx(row["CL_me"], rel=1e-15), k
        if np.isnan(row["CL"]):
            assert np.isnan(r["series_CL"][k])
        else:
            assert r["series_CL"][k] == pytest.approx(row["CL"], rel=1e-12)
            assert r["series_CD"][k] == pytest.approx(row["CD"], rel=1e-12)
        assert r["series_sep_frac"][k] == pytest.approx(row["sep"], rel=1e-12, abs=1e-15)
    assert_bitwise(r["F"], o.F, "gloo slabs vs whole lattice")
    assert tuple(r["me"]) == tuple(o.me_hist[-1])
    assert float(r["mass"][0]) == pytest.approx(olbm.total_mass(o.F), rel=1e-13)
    assert float(r["tmax"]) == world and float(r["mn"][0]) == 0.0
    assert bytes(r["blobs"]) == b"".join(bytes([k]) * 4 for k in range(world))

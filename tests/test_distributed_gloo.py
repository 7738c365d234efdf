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
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_worker.py"), str(nx), str(ny),
                               str(nsteps), out], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
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
    r = np.load(out)
    assert_bitwise(r["F"], o.F, "gloo slabs vs whole lattice")
    assert tuple(r["me"]) == tuple(o.me_hist[-1])
    assert float(r["mass"][0]) == pytest.approx(olbm.total_mass(o.F), rel=1e-13)
    assert float(r["tmax"]) == world and float(r["mn"][0]) == 0.0
    assert bytes(r["blobs"]) == b"".join(bytes([k]) * 4 for k in range(world))

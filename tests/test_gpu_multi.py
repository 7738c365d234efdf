"""GPU tests that need two devices: slabs on different GPUs with the in-kernel NVLink halo push."""
import numpy as np
import pytest

from conftest import assert_bitwise

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def al(built_lib):
    import aerolab_lbm
    if aerolab_lbm.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    return aerolab_lbm


@pytest.mark.parametrize("nx,ny,double", [(512, 256, 0), (1400, 256, 1)])
def test_two_device_slabs_bitwise(al, nx, ny, double):
    """double=1: two steps per pass; the halo is pushed by the list-driven passes over NVLink."""
    whole = al.WindTunnel(nx, ny, 0)
    whole.load_shape("naca2412", alpha=7.0)
    a = al.WindTunnel(nx, ny, 0, y0=0, ny_local=120)
    b = al.WindTunnel(nx, ny, 1, y0=120, ny_local=136)
    for s in (a, b):
        s.set_double_steps(double)
        s.load_shape("naca2412", alpha=7.0)
    a.connect_local(None, b)
    b.connect_local(a, None)
    n = 80
    whole.step(n)
    # several steps per call: the stream-ordered flag kernels order the two GPUs
    for _ in range(n // 8):
        a.step(8)
        b.step(8)
    a.sync(); b.sync()
    assert_bitwise(np.concatenate([a.populations(), b.populations()], 1), whole.populations(), "2-GPU populations")
    wm = whole.macro()
    am, bm = a.macro(), b.macro()
    for k in range(3):
        assert_bitwise(np.concatenate([am[k], bm[k]], 0), wm[k], f"2-GPU macro {k}")
    me = a.me_history(1)[0] + b.me_history(1)[0]
    assert np.array_equal(me, whole.me_history(1)[0])


@pytest.mark.parametrize("halo", ["p2p", "nccl", "p2p-double", "p2p-balanced-double"])
def test_distributed_tunnel_torchrun(al, halo):
    """One process per GPU (torchrun, NCCL plumbing): IPC/NVLink halo push and the NCCL fallback
    must both reproduce the single-GPU run bit for bit."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    n = min(al.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", {"p2p": "29577", "nccl": "29578", "p2p-double": "29579"}.get(halo, "29580"),
           os.path.join(ROOT, "tests", "_dist_gpu_worker.py"), halo]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "bitwise_ok=True" in r.stdout


def test_create_multi_across_devices(al):
    n = min(al.device_count(), 4)
    nx, ny = 1024, 400
    whole = al.WindTunnel(nx, ny, 0)
    whole.load_shape("naca4412", alpha=6.0)
    multi = al.LocalMultiTunnel(nx, ny, list(range(n)))
    multi.load_shape("naca4412", alpha=6.0)
    whole.step(64); multi.step(64); multi.sync()
    assert_bitwise(multi.populations(), whole.populations(), "multi-GPU populations")
    multi.close()

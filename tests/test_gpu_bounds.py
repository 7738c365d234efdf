"""GPU: run the step kernels of a bounds-checked debug build (ALB_DEBUG_BOUNDS=1) over awkward
lattices.  compute-sanitizer is not available on the GPU pool, so the library carries its own
check: every population load/store address is validated against its allocation and a violation
traps (surfacing as a CUDA error from alb_sync).  Runs in a subprocess because the debug build is
a different shared library (AEROLAB_LBM_LIB)."""
import os
import subprocess
import sys

import pytest

from conftest import PKG_DIR, ROOT

pytestmark = pytest.mark.gpu

SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import aerolab_lbm as al
rng = np.random.default_rng(0)
for nx, ny in ((3, 3), (5, 4), (127, 9), (128, 8), (129, 7), (320, 160), (513, 33), (1024, 300)):
    t = al.WindTunnel(nx, ny, 0)
    m = (rng.random((ny, nx)) < 0.15).astype(np.uint8) * 255
    t.set_mask(m)
    for n in (1, 1, 7, 2):          # single-launch, persistent and split paths
        t.step(n)
    t.macro(); t.forces(); t.update_stats(); t.sync()
    t.close()
# slabs with halo pushes
a = al.WindTunnel(200, 50, 0, y0=0, ny_local=1); b = al.WindTunnel(200, 50, 0, y0=1, ny_local=49)
for s in (a, b): s.load_shape("naca0012", alpha=4.0)
a.connect_local(None, b); b.connect_local(a, None)
for _ in range(10):
    a.step(1); b.step(1)
a.sync(); b.sync()
big = al.WindTunnel(4096, 600, 0); big.load_shape("naca4412", alpha=10.0); big.step(3); big.forces(); big.sync()
# two steps per pass: fused kernel (strip margins, last strip, short last segment), list-driven passes, solid copies
for nx, ny in ((320, 160), (641, 131), (1279, 5), (1281, 140), (2500, 260), (130, 3)):
    t = al.WindTunnel(nx, ny, 0)
    t.set_double_steps(1)
    t.load_shape("naca4412", alpha=10.0)
    for n in (3, 4, 11, 1, 5):
        t.step(n)
    t.macro(); t.forces(); t.update_stats(); t.sync()
    t.close()
a = al.WindTunnel(700, 90, 0, y0=0, ny_local=40); b = al.WindTunnel(700, 90, 0, y0=40, ny_local=50)
for s in (a, b): s.set_double_steps(1); s.load_shape("naca0012", alpha=4.0)
a.connect_local(None, b); b.connect_local(a, None)
for _ in range(3):
    a.step(5); b.step(5)
a.sync(); b.sync()
print("BOUNDS_OK")
""" % PKG_DIR


def test_bounds_checked_build_runs_clean(tmp_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("alb_build", os.path.join(PKG_DIR, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = os.path.join(ROOT, "variants", "lib_bounds.so")
    srcs = [os.path.join(mod.CSRC, d) for d in mod.DEPS]
    stale = not os.path.exists(lib) or any(os.path.getmtime(p) > os.path.getmtime(lib) for p in srcs)
    if stale:
        try:
            mod.nvcc_path()
        except RuntimeError:
            if not os.path.exists(lib):
                pytest.skip("no nvcc and no prebuilt variants/lib_bounds.so")
        else:
            mod.build(defines=["ALB_DEBUG_BOUNDS=1"], out=lib)
    r = subprocess.run([sys.executable, "-c", SCRIPT], env=dict(os.environ, AEROLAB_LBM_LIB=lib),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "BOUNDS_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bounds violation" not in r.stdout

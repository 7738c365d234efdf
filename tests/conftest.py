"""pytest configuration: the `gpu` marker, import paths, shared helpers."""
import os
import sys

import numpy as np
import pytest

# In-process slabs that share ONE device wait for each other inside kernels; streams that alias onto
# the same hardware queue could then block each other (one process per GPU, the product layout, has
# no such coupling).  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "airfoil-cfd-tool_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu() -> bool:
    try:
        from aerolab_lbm import device_count
        return device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip silently; plain
    # runs without a GPU skip the gpu tests.
    if _have_gpu():
        return
    selected_gpu = "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or "")
    if selected_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a.view(np.uint64)


def assert_bitwise(a, b, what=""):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    same = bits(a) == bits(b)
    if not same.all():
        idx = np.argwhere(~same)
        first = tuple(idx[0])
        raise AssertionError(f"{what}: {len(idx)} of {same.size} values differ bitwise; first at {first}: "
                             f"{a[first]!r} vs {b[first]!r}")


def rel_linf(a, ref, fluid=None):
    """L-infinity error normalised by max |ref| (over fluid cells if a mask is given)."""
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    if fluid is not None:
        a, ref = a[..., fluid], ref[..., fluid]
    return float(np.max(np.abs(a - ref)) / np.max(np.abs(ref)))


@pytest.fixture(scope="session")
def built_lib():
    """Build libaerolab_lbm.so when a toolchain is present and the binary is stale."""
    sys.path.insert(0, PKG_DIR)
    import importlib.util
    spec = importlib.util.spec_from_file_location("alb_build", os.path.join(PKG_DIR, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        mod.nvcc_path()
        mod.build()
    except RuntimeError:
        pass
    from aerolab_lbm import _ffi
    return _ffi.lib()

"""CPU: the C-ABI library loads and exports exactly what include/aerolab_lbm.h declares."""
import os
import re
import subprocess

import pytest

from conftest import ROOT


def header_functions():
    text = open(os.path.join(ROOT, "include", "aerolab_lbm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(alb_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from aerolab_lbm import _ffi
    assert header_functions() == sorted(_ffi.EXPORTS)


def test_library_exports_every_declared_symbol(built_lib):
    from aerolab_lbm import _ffi
    for name in header_functions():
        assert hasattr(built_lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (alb_[a-z0-9_]+)", out))
    assert set(header_functions()) <= exported
    assert built_lib.alb_version() == 100
    assert built_lib.alb_error_string(-5) == b"halo wait timed out"


def test_no_silent_cpu_fallback(built_lib):
    """Without a CUDA device creation must fail with ALB_ERR_CUDA, never compute on the CPU."""
    import aerolab_lbm as al
    if al.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(al.AerolabLbmError) as e:
        al.WindTunnel(64, 32, 0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_missing_library_fails_loudly(monkeypatch):
    from aerolab_lbm import _ffi
    monkeypatch.setattr(_ffi, "_lib", None)
    monkeypatch.setattr(_ffi, "LIB_PATH", "/nonexistent/libaerolab_lbm.so")
    with pytest.raises(_ffi.AerolabLbmError):
        _ffi.lib()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "airfoil-cfd-tool_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liblbm_oracle" not in text, f

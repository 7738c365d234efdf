"""ctypes binding of libaerolab_lbm.so (C ABI: include/aerolab_lbm.h).

There is no CPU fallback.  If the shared library has not been built, or no CUDA
device is present when a tunnel is created, this module raises
``AerolabLbmError`` -- it never routes around the CUDA path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# AEROLAB_LBM_LIB: use another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("AEROLAB_LBM_LIB") or os.path.join(_HERE, "_lib", "libaerolab_lbm.so")

ALB_OK = 0
ALB_ERR_INVALID = -1
ALB_ERR_CUDA = -2
ALB_ERR_NOMEM = -3
ALB_ERR_STATE = -4
ALB_ERR_TIMEOUT = -5
ALB_NPANEL = 160
ALB_IPC_BYTES = 256
ALB_ME_HISTORY = 4095
ALB_ME_SCALE = float(2 ** 40)
ALB_FRAME_ROW = 12

# every symbol include/aerolab_lbm.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "alb_version", "alb_error_string", "alb_last_error", "alb_device_count",
    "alb_create", "alb_create_slab", "alb_destroy", "alb_get_dims",
    "alb_set_params", "alb_get_params", "alb_reset",
    "alb_rasterize", "alb_rasterize_panels", "alb_set_mask", "alb_get_mask", "alb_get_panels",
    "alb_step", "alb_sync", "alb_step_count", "alb_last_step_ms",
    "alb_get_populations", "alb_get_population_rows", "alb_state_hash", "alb_set_populations", "alb_get_macro", "alb_set_macro", "alb_total_mass",
    "alb_update_stats", "alb_stats_partial", "alb_set_stats", "alb_get_stats",
    "alb_get_field", "alb_get_rgba", "alb_get_macro_edges", "alb_set_macro_ghosts",
    "alb_compute_forces", "alb_forces_partial", "alb_reset_force_emas",
    "alb_get_me_history", "alb_get_me_forces", "alb_clamp_hits",
    "alb_reynolds", "alb_stall_state",
    "alb_run_frames", "alb_frames_enqueue", "alb_frames_collect",
    "alb_particles_init", "alb_particles_resize", "alb_particles_step", "alb_particles_get",
    "alb_create_multi", "alb_step_multi", "alb_destroy_multi", "alb_connect_local", "alb_ipc_export", "alb_ipc_connect", "alb_halo_prime", "alb_halo_ptrs",
    "alb_set_external_halo", "alb_set_double_steps", "alb_selftest_division", "alb_launch_count", "alb_get_double_steps", "alb_debug_step2_plan", "alb_get_div_mode", "alb_set_div_mode",
]


class AerolabLbmError(RuntimeError):
    """Raised for every non-zero return code of the C ABI."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code
        self.message = message


_lib = None


def lib():
    """Load the shared library (once).  Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AerolabLbmError(
            ALB_ERR_STATE,
            f"{LIB_PATH} not found: build it with `python airfoil-cfd-tool_b200/build.py` "
            "(needs nvcc; the library targets sm_100a and has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    vp = C.c_void_p
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int)
    L.alb_version.restype = C.c_int
    L.alb_error_string.restype = C.c_char_p
    L.alb_error_string.argtypes = [C.c_int]
    L.alb_last_error.restype = C.c_char_p
    L.alb_last_error.argtypes = [H]
    L.alb_device_count.argtypes = [ip]
    L.alb_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(H)]
    L.alb_create_slab.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(H)]
    L.alb_destroy.argtypes = [H]
    L.alb_get_dims.argtypes = [H, ip, ip, ip, ip]
    L.alb_set_params.argtypes = [H, C.c_double, C.c_double]
    L.alb_get_params.argtypes = [H, dp, dp]
    L.alb_reset.argtypes = [H, C.c_double]
    L.alb_rasterize.argtypes = [H, vp, C.c_int, C.c_double, vp]
    L.alb_rasterize_panels.argtypes = [H, vp, vp, C.c_int, vp]
    L.alb_set_mask.argtypes = [H, vp]
    L.alb_get_mask.argtypes = [H, vp]
    L.alb_get_panels.argtypes = [H, vp, vp]
    L.alb_step.argtypes = [H, C.c_int]
    L.alb_sync.argtypes = [H]
    L.alb_step_count.argtypes = [H, C.POINTER(C.c_longlong)]
    L.alb_last_step_ms.argtypes = [H, C.POINTER(C.c_float)]
    L.alb_get_populations.argtypes = [H, vp]
    L.alb_set_populations.argtypes = [H, vp]
    L.alb_get_population_rows.argtypes = [H, C.c_int, C.c_int, vp]
    L.alb_state_hash.argtypes = [H, vp]
    L.alb_get_macro.argtypes = [H, vp, vp, vp]
    L.alb_set_macro.argtypes = [H, vp, vp, vp]
    L.alb_total_mass.argtypes = [H, dp]
    L.alb_update_stats.argtypes = [H, vp, vp, vp, vp]
    L.alb_stats_partial.argtypes = [H, vp, vp, vp, vp]
    L.alb_set_stats.argtypes = [H, C.c_double, C.c_double, C.c_double]
    L.alb_get_stats.argtypes = [H, vp]
    L.alb_get_field.argtypes = [H, C.c_int, vp]
    L.alb_get_rgba.argtypes = [H, C.c_int, vp]
    L.alb_get_macro_edges.argtypes = [H, vp, vp]
    L.alb_set_macro_ghosts.argtypes = [H, vp, vp]
    L.alb_compute_forces.argtypes = [H, vp]
    L.alb_forces_partial.argtypes = [H, vp]
    L.alb_reset_force_emas.argtypes = [H]
    L.alb_get_me_history.argtypes = [H, C.c_int, vp]
    L.alb_get_me_forces.argtypes = [H, vp]
    L.alb_clamp_hits.argtypes = [H, C.POINTER(C.c_longlong)]
    L.alb_reynolds.argtypes = [H, dp]
    L.alb_stall_state.argtypes = [H, ip, ip]
    L.alb_run_frames.argtypes = [H, C.c_int, C.c_int, C.c_int, vp, vp]
    L.alb_frames_enqueue.argtypes = [H, C.c_int, C.c_int, C.c_int, vp]
    L.alb_frames_collect.argtypes = [H, vp]
    L.alb_particles_init.argtypes = [H, C.c_int, C.c_ulonglong]
    L.alb_particles_resize.argtypes = [H, C.c_int]
    L.alb_particles_step.argtypes = [H, C.c_double]
    L.alb_particles_get.argtypes = [H, vp, ip]
    L.alb_create_multi.argtypes = [C.c_int, C.c_int, ip, C.c_int, C.POINTER(H)]
    L.alb_step_multi.argtypes = [C.POINTER(H), C.c_int, C.c_int]
    L.alb_destroy_multi.argtypes = [C.POINTER(H), C.c_int]
    L.alb_connect_local.argtypes = [H, H, H]
    L.alb_ipc_export.argtypes = [H, vp]
    L.alb_ipc_connect.argtypes = [H, vp, vp]
    L.alb_halo_prime.argtypes = [H]
    L.alb_halo_ptrs.argtypes = [H, vp, vp, vp, vp]
    L.alb_set_external_halo.argtypes = [H, C.c_int]
    L.alb_set_double_steps.argtypes = [H, C.c_int]
    L.alb_launch_count.argtypes = [H, C.POINTER(C.c_longlong)]
    L.alb_get_double_steps.argtypes = [H, ip, ip]
    L.alb_debug_step2_plan.argtypes = [C.c_int, C.c_int, C.c_int, ip]
    L.alb_get_div_mode.argtypes = [H, ip]
    L.alb_set_div_mode.argtypes = [H, C.c_int]
    L.alb_selftest_division.argtypes = [H, C.c_ulonglong, C.c_longlong, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("alb_error_string", "alb_last_error"):
            fn.restype = C.c_int
    _lib = L
    return L


def check(code: int, handle=None) -> None:
    if code == ALB_OK:
        return
    L = lib()
    msg = L.alb_last_error(handle)
    text = msg.decode("utf-8", "replace") if msg else ""
    if not text:
        text = L.alb_error_string(code).decode()
    raise AerolabLbmError(code, text)


def ptr(a):
    """void* of a C-contiguous NumPy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def device_count() -> int:
    n = C.c_int(0)
    code = lib().alb_device_count(C.byref(n))
    if code != ALB_OK:
        return 0
    return n.value

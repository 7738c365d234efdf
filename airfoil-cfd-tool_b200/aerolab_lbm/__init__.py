"""aerolab_lbm -- B200-native D2Q9 lattice-Boltzmann airfoil wind tunnel.

Python host of libaerolab_lbm.so (C ABI in include/aerolab_lbm.h).  It keeps
the control surface of the reference's interactive tunnel
(pages/airfoil_flow_lbm_aerolab.html): angle of attack, inlet velocity,
relaxation time / viscosity, speed / Cp / vorticity field modes, CL/CD and the
stall indicator.  CUDA only: no Triton, no backend dispatch, no CPU fallback.
"""
from ._ffi import AerolabLbmError, LIB_PATH, device_count
from .geometry import SHAPES, clark_y, naca4, naca_digits, round_coords
from .tunnel import (DEFAULT_ALPHA, DEFAULT_NX, DEFAULT_NY, DEFAULT_TAU, DEFAULT_U0, LocalMultiTunnel,
                     WindTunnel, build_lbm_component, state_hash_numpy)

__all__ = [
    "AerolabLbmError", "LIB_PATH", "device_count", "SHAPES", "clark_y", "naca4", "naca_digits",
    "round_coords", "WindTunnel", "LocalMultiTunnel", "build_lbm_component", "state_hash_numpy", "DEFAULT_ALPHA", "DEFAULT_NX", "DEFAULT_NY",
    "DEFAULT_TAU", "DEFAULT_U0",
]
__version__ = "0.1.0"

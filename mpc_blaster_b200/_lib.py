"""ctypes loader for libmpcb.so (the CUDA library behind include/mpcb.h).

There is deliberately no CPU fallback: if the library is missing or no CUDA device is
present, creating a solver raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

MPCB_SHARED, MPCB_PER_INSTANCE, MPCB_PER_STAGE = 0, 1, 2


class MpcbConfig(C.Structure):
    """Mirror of ``mpcb_config`` (include/mpcb.h)."""
    _fields_ = [("variant", C.c_int32), ("N", C.c_int32), ("dt", C.c_double), ("mass", C.c_double),
                ("J", C.c_double * 9), ("l_x", C.c_double), ("l_y", C.c_double), ("c", C.c_double),
                ("Q", C.c_double * 17), ("R", C.c_double * 6), ("Qt", C.c_double * 17),
                ("lbx", C.c_double * 17), ("ubx", C.c_double * 17), ("lbu", C.c_double * 6), ("ubu", C.c_double * 6),
                ("ipm_max_iter", C.c_int32), ("ipm_mu0", C.c_double), ("ipm_thr0", C.c_double),
                ("tol_stat", C.c_double), ("tol_eq", C.c_double), ("tol_ineq", C.c_double), ("tol_comp", C.c_double),
                ("alpha_min", C.c_double), ("dtype", C.c_int32), ("max_batch", C.c_int32), ("ws_batch", C.c_int32), ("device", C.c_int32),
                ("strict_reference", C.c_int32), ("throughput_batch", C.c_int32), ("qp8_batch", C.c_int32), ("qp8_warps", C.c_int32)]


# every symbol include/mpcb.h declares
EXPORTS = ["mpcb_config_default", "mpcb_create", "mpcb_destroy", "mpcb_last_error", "mpcb_nx", "mpcb_nu", "mpcb_horizon",
           "mpcb_reset", "mpcb_solve", "mpcb_solve_host", "mpcb_plant_step", "mpcb_closed_loop", "mpcb_cost",
           "mpcb_get_iterate", "mpcb_set_iterate", "mpcb_debug_linearize", "mpcb_kernel_launches", "mpcb_command_map",
           "mpcb_profile", "mpcb_last_kernel_ms", "mpcb_fp64_peak", "mpcb_solve_sqp", "mpcb_shift", "mpcb_poc_jacobians",
           "mpcb_debug_qp"]

_lib = None


def library_path() -> str:
    return _build.LIB


def load() -> C.CDLL:
    """Load (building first if the sources are newer and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MPCB_LIB_OVERRIDE") or _build.LIB  # override: A/B builds of tools/ab.py only
    if path == _build.LIB and _build._stale():
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        if os.path.exists(nvcc):
            _build.build()
        elif not os.path.exists(path):
            raise RuntimeError(f"{path} is missing and nvcc is not available: the solver has no CPU fallback; "
                               "run `python -m mpc_blaster_b200._build` where nvcc exists")
    lib = C.CDLL(path)
    vp, dp, ip = C.c_void_p, C.c_void_p, C.c_void_p  # raw device/host addresses
    lib.mpcb_config_default.argtypes = [C.POINTER(MpcbConfig), C.c_int, C.c_int]
    lib.mpcb_create.argtypes = [C.POINTER(MpcbConfig), C.POINTER(vp)]
    lib.mpcb_destroy.argtypes = [vp]
    lib.mpcb_last_error.argtypes = [vp]
    lib.mpcb_last_error.restype = C.c_char_p
    for f in ("mpcb_nx", "mpcb_nu", "mpcb_horizon"):
        getattr(lib, f).argtypes = [vp]
    lib.mpcb_reset.argtypes = [vp, dp, dp, C.c_int, C.c_int, vp]
    lib.mpcb_solve.argtypes = [vp, dp, dp, C.c_int, dp, C.c_int, dp, dp, dp, ip, ip, C.c_int, vp]
    lib.mpcb_solve_sqp.argtypes = [vp, dp, dp, C.c_int, dp, C.c_int, C.c_int, dp, dp, dp, dp, ip, ip, ip, dp, C.c_int, vp]
    lib.mpcb_debug_qp.argtypes = [vp] + [dp] * 11 + [C.c_int, vp]
    lib.mpcb_shift.argtypes = [vp, C.c_int, vp]
    lib.mpcb_solve_host.argtypes = [vp, dp, dp, C.c_int, dp, C.c_int, dp, dp, dp, ip, ip, C.c_int]
    lib.mpcb_plant_step.argtypes = [vp, dp, dp, dp, C.c_int, dp, C.c_int, vp]
    lib.mpcb_closed_loop.argtypes = [vp, dp, dp, C.c_int, dp, C.c_int, C.c_int, dp, ip, ip, C.c_int, vp]
    lib.mpcb_cost.argtypes = [vp, dp, C.c_int, dp, C.c_int, vp]
    lib.mpcb_get_iterate.argtypes = [vp, dp, dp, C.c_int, vp]
    lib.mpcb_set_iterate.argtypes = [vp, dp, dp, C.c_int, vp]
    lib.mpcb_debug_linearize.argtypes = [vp, dp, C.c_int, dp, dp, C.c_int, vp]
    lib.mpcb_command_map.argtypes = [vp, dp, dp, dp, dp, C.c_int, vp]
    lib.mpcb_poc_jacobians.argtypes = [dp, dp, dp, dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, dp, dp, dp, dp, dp, dp, ip, vp]
    lib.mpcb_kernel_launches.restype = C.c_int64
    lib.mpcb_profile.argtypes = [vp, C.c_int]
    lib.mpcb_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.mpcb_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    _lib = lib
    return lib

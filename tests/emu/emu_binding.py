"""ctypes binding of the host-emulated kernel bodies (tests/emu/emu_main.cpp).  Test-only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "mpc_blaster_b200", "csrc")
# MPCB_EMU_DEFINES="A,B=1": host build of an experimental -D variant of the kernels (tools/ab.py's CPU counterpart)
_DEFS = [d for d in os.environ.get("MPCB_EMU_DEFINES", "").split(",") if d]
LIB = os.path.join(HERE, "_build", "libmpcb_emu" + "".join("_" + d.replace("=", "") for d in _DEFS) + ".so")


class Params(C.Structure):
    """Mirror of mpcb::Params (mpc_blaster_b200/csrc/mpcb_common.cuh)."""
    _fields_ = [("variant", C.c_int), ("N", C.c_int), ("dt", C.c_double), ("mass", C.c_double), ("inv_mass", C.c_double),
                ("J", C.c_double * 9), ("Jinv", C.c_double * 9), ("JinvG", C.c_double * 12),
                ("l_x", C.c_double), ("l_y", C.c_double), ("c", C.c_double),
                ("Q", C.c_double * 17), ("R", C.c_double * 6), ("Qt", C.c_double * 17),
                ("lbx", C.c_double * 17), ("ubx", C.c_double * 17), ("lbu", C.c_double * 6), ("ubu", C.c_double * 6),
                ("ipm_max_iter", C.c_int), ("ipm_mu0", C.c_double), ("ipm_thr0", C.c_double),
                ("tol_stat", C.c_double), ("tol_eq", C.c_double), ("tol_ineq", C.c_double), ("tol_comp", C.c_double),
                ("alpha_min", C.c_double), ("strict", C.c_int), ("reserved_", C.c_int)]


def build(force=False):
    srcs = [os.path.join(HERE, "emu_main.cpp"), os.path.join(HERE, "warp_emu.h")] + \
           [os.path.join(CSRC, f) for f in ("mpcb_common.cuh", "mpcb_model.cuh", "mpcb_linearize.cuh", "mpcb_qp.cuh", "mpcb_qp8.cuh", "mpcb_poc.cuh")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-DMPCB_HOST_EMU", "-ffp-contract=off", "-fPIC", "-shared"] + ["-D" + d for d in _DEFS] + [
                           "-I" + HERE, "-I" + CSRC, "-x", "c++", srcs[0], "-o", LIB])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.emu_params_size.restype = C.c_size_t
        assert _lib.emu_params_size() == C.sizeof(Params)
    return _lib


def make_params(P, max_iter=None, mu0=1e2, thr0=-0.5, tol_stat=1e-6, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8,
                alpha_min=1e-8, strict=False) -> Params:
    """P: oracle.blaster_oracle.BlasterProblem (only used here as a container of constants)."""
    o = Params()
    if max_iter is None:
        max_iter = 500 if strict else 60
    o.strict = int(bool(strict))
    o.variant, o.N, o.dt, o.mass, o.inv_mass = P.variant, P.N, P.dt, P.mass, 1.0 / P.mass
    o.J[:] = P.J.reshape(-1)
    Jinv = np.linalg.inv(P.J)
    o.Jinv[:] = Jinv.reshape(-1)
    G = np.array([[-P.l_y, P.l_y, -P.l_y, P.l_y], [-P.l_x, P.l_x, P.l_x, -P.l_x], [-P.c, -P.c, P.c, P.c]])
    o.JinvG[:] = (Jinv @ G).reshape(-1)
    o.l_x, o.l_y, o.c = P.l_x, P.l_y, P.c
    for name, n in (("Q", P.nx), ("R", P.nu), ("Qt", P.nx), ("lbx", P.nx), ("ubx", P.nx), ("lbu", P.nu), ("ubu", P.nu)):
        arr, v = getattr(o, name), getattr(P, name)
        for i in range(n):
            arr[i] = v[i]
    o.ipm_max_iter, o.ipm_mu0, o.ipm_thr0 = max_iter, mu0, thr0
    o.tol_stat, o.tol_eq, o.tol_ineq, o.tol_comp, o.alpha_min = tol_stat, tol_eq, tol_ineq, tol_comp, alpha_min
    return o


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def rti_solve(P, X, U, x0, yref, p, **opts):
    """One emulated RTI iteration on one instance.  X, U are updated in place."""
    o = make_params(P, **opts)
    yref = np.ascontiguousarray(yref, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    nz = P.nx + P.nu
    BAt = np.zeros((P.N, nz, P.nx))
    b = np.zeros((P.N, P.nx))
    it = C.c_int(0)
    st = lib().emu_rti_solve(C.byref(o), _dp(X), _dp(U), _dp(x0), _dp(yref), int(yref.ndim == 2), _dp(p), int(p.ndim == 2),
                             C.byref(it), _dp(BAt), _dp(b))
    return st, it.value, BAt, b


def plant_step(P, x, u, p):
    o = make_params(P)
    xn = np.zeros(P.nx)
    lib().emu_plant_step(C.byref(o), _dp(np.ascontiguousarray(x, dtype=np.float64)), _dp(np.ascontiguousarray(u, dtype=np.float64)),
                         _dp(np.ascontiguousarray(p, dtype=np.float64)), _dp(xn))
    return xn


def poc(euler, motor, position, V=150.0, drag=1.0, mode=0):
    """-> (poc[3], J_mot[3,2], J_eul[3,3], J_pos[3,3], t_flight, status, p25)"""
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    e, m, p = (np.ascontiguousarray(a, dtype=np.float64) for a in (euler, motor, position))
    out, p25 = np.zeros(29), np.zeros(25)
    f = lib().emu_poc
    f.argtypes = [C.POINTER(C.c_double)] * 3 + [C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    f.restype = None
    f(dp(e), dp(m), dp(p), V, drag, mode, dp(out), dp(p25))
    J = out[3:27].reshape(3, 8)
    return out[:3].copy(), J[:, 0:2].copy(), J[:, 2:5].copy(), J[:, 5:8].copy(), out[27], int(out[28]), p25


def rti_solve4(P, X, U, x0, yref, p, **opts):
    """One emulated RTI iteration of nb instances by ONE warp of the four-instances-per-warp kernel
    (mpcb_qp8.cuh): its four groups draw the instances from a work counter and are refilled as they finish.
    X[nb,N+1,nx], U[nb,N,nu] are updated in place; returns (status[nb], iters[nb])."""
    o = make_params(P, **opts)
    nb = X.shape[0]
    assert nb >= 1 and X.flags.c_contiguous and U.flags.c_contiguous
    x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(nb, P.nx)
    yref = np.ascontiguousarray(yref, dtype=np.float64).reshape(nb, P.nx + P.nu)
    p = np.ascontiguousarray(p, dtype=np.float64)
    st, it = np.zeros(nb, dtype=np.int32), np.zeros(nb, dtype=np.int32)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    lib().emu_rti_solve4(C.byref(o), nb, _dp(X), _dp(U), _dp(x0), _dp(yref), _dp(p), ip(st), ip(it))
    return st, it

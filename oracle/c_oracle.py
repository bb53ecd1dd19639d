"""ctypes binding of the C oracle (oracle/mpc_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .blaster_oracle import BlasterProblem, default_params
from .build_oracle import LIB, build


class OrcProblem(C.Structure):
    _fields_ = [("variant", C.c_int), ("N", C.c_int), ("dt", C.c_double), ("mass", C.c_double),
                ("J", C.c_double * 9), ("Jinv", C.c_double * 9), ("l_x", C.c_double), ("l_y", C.c_double),
                ("c", C.c_double), ("Q", C.c_double * 17), ("R", C.c_double * 6), ("Qt", C.c_double * 17),
                ("lbx", C.c_double * 17), ("ubx", C.c_double * 17), ("lbu", C.c_double * 6), ("ubu", C.c_double * 6),
                ("ipm_max_iter", C.c_int), ("rg_mode", C.c_int), ("ric_alg", C.c_int), ("ipm_mu0", C.c_double), ("ipm_thr0", C.c_double),
                ("tol_stat", C.c_double), ("tol_eq", C.c_double), ("tol_ineq", C.c_double), ("tol_comp", C.c_double), ("alpha_min", C.c_double),
                ("strict", C.c_int), ("itref", C.c_int), ("mixed_mu", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = LIB if os.path.exists(LIB) and not os.path.exists("/root/reference") else build()
        _lib = C.CDLL(path)
        assert _lib.orc_problem_size() == C.sizeof(OrcProblem)
        _lib.orc_rti_solve_batch.restype = C.c_int
        _lib.orc_sqp_solve_batch.restype = C.c_int
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_problem(P: BlasterProblem, max_iter=None, mu0=1e2, thr0=-0.5, tol_stat=1e-6, tol_eq=1e-8, tol_ineq=1e-8,
                 tol_comp=1e-8, ric_alg=1, rg_mode=2, alpha_min=1e-8, strict=False, itref=None, mixed_mu=0.0) -> OrcProblem:
    """``strict``: the reference stack's semantics (mpcb_config.strict_reference) -- explicit residual norms in the stopping
    test, no divergence exit, last iterate applied on max-iter, iteration cap 500 (blastermodel.py:279) unless given, one
    step of iterative refinement on the corrector solve (what makes the explicit stationarity norm reach 1e-6)."""
    o = OrcProblem()
    if max_iter is None:
        max_iter = 500 if strict else 60
    o.strict = int(bool(strict))
    o.mixed_mu = float(mixed_mu)
    o.itref = int((1 if strict else 0) if itref is None else itref)  # strict: one refinement step on the corrector solve, as the kernel's STRICT instantiation
    o.variant, o.N, o.dt, o.mass = P.variant, P.N, P.dt, P.mass
    o.J[:] = P.J.reshape(-1)
    o.Jinv[:] = P.Jinv.reshape(-1)
    o.l_x, o.l_y, o.c = P.l_x, P.l_y, P.c
    for name, n in (("Q", P.nx), ("R", P.nu), ("Qt", P.nx), ("lbx", P.nx), ("ubx", P.nx), ("lbu", P.nu), ("ubu", P.nu)):
        arr = getattr(o, name)
        v = getattr(P, name)
        for i in range(n):
            arr[i] = v[i]
    o.ipm_max_iter, o.ipm_mu0, o.ipm_thr0, o.ric_alg = max_iter, mu0, thr0, ric_alg
    o.rg_mode = rg_mode
    o.alpha_min = alpha_min
    o.tol_stat, o.tol_eq, o.tol_ineq, o.tol_comp = tol_stat, tol_eq, tol_ineq, tol_comp
    return o


def f(P: BlasterProblem, x, u, p):
    o = make_problem(P)
    x = np.ascontiguousarray(x, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    out = np.zeros(P.nx)
    lib().orc_f(C.byref(o), _dp(x), _dp(u), _dp(p), _dp(out))
    return out


def rk4_sens(P: BlasterProblem, x, u, p):
    o = make_problem(P)
    x = np.ascontiguousarray(x, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    xn = np.zeros(P.nx)
    BAt = np.zeros((P.nx + P.nu, P.nx))
    lib().orc_rk4_sens(C.byref(o), _dp(x), _dp(u), _dp(p), _dp(xn), _dp(BAt))
    return xn, BAt[P.nu:].T.copy(), BAt[:P.nu].T.copy()


def plant_step(P: BlasterProblem, x, u, p=None, nthreads=1):
    o = make_problem(P)
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, P.nx)
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1, P.nu)
    p = default_params() if p is None else np.ascontiguousarray(p, dtype=np.float64)
    xn = np.zeros_like(x)
    lib().orc_plant_step_batch(C.byref(o), _dp(x), _dp(u), _dp(p), int(p.ndim == 2), _dp(xn), x.shape[0], nthreads)
    return xn


class BatchRTI:
    """B independent controllers with persistent un-shifted iterate, OpenMP over instances."""

    def __init__(self, P: BlasterProblem, B: int, nthreads: int | None = None, **opts):
        self.P, self.B = P, B
        self.o = make_problem(P, **opts)
        self.nthreads = nthreads or lib().orc_max_threads()
        self.reset()

    def reset(self, x_init=None, u_init=None):
        """[upstream D4] zero iterate unless the caller sets one (acados: set(k,'x'/'u',...)).
        x_init[B,nx] / u_init[nu] or [B,nu] are broadcast over the stages."""
        P = self.P
        self.X = np.zeros((self.B, P.N + 1, P.nx))
        self.U = np.zeros((self.B, P.N, P.nu))
        if x_init is not None:
            self.X[:] = np.asarray(x_init, dtype=np.float64).reshape(self.B, 1, P.nx)
        if u_init is not None:
            self.U[:] = np.asarray(u_init, dtype=np.float64).reshape(-1, 1, P.nu)
        self.status = np.zeros(self.B, dtype=np.int32)
        self.iters = np.zeros(self.B, dtype=np.int32)

    def solve(self, x0, yref, p=None):
        P, B = self.P, self.B
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(B, P.nx)
        yref = np.ascontiguousarray(yref, dtype=np.float64)
        ymode = {1: 0, 2: 1, 3: 2}[yref.ndim]
        p = default_params() if p is None else np.ascontiguousarray(p, dtype=np.float64)
        pmode = {1: 0, 2: 1, 3: 2}[p.ndim]
        if ymode == 2:
            assert yref.shape == (B, P.N + 1, P.ny)
        if pmode == 2:
            assert p.shape == (B, P.N, 25)
        rc = lib().orc_rti_solve_batch(C.byref(self.o), _dp(self.X), _dp(self.U), _dp(x0), _dp(yref), ymode, _dp(p), pmode,
                                       self.status.ctypes.data_as(C.POINTER(C.c_int32)),
                                       self.iters.ctypes.data_as(C.POINTER(C.c_int32)), B, self.nthreads)
        assert rc == 0
        return self.U[:, 0].copy(), self.X.copy(), self.U.copy(), self.status.copy()

    def sqp_solve(self, x0, yref, p=None, max_iter=100, tol=1e-6):
        """SQP to convergence: -> (u0, X, U, status, sqp_iters, qp_iters, res[B,4]).
        tol: a float or (stat, eq, ineq, comp)."""
        P, B = self.P, self.B
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(B, P.nx)
        yref = np.ascontiguousarray(yref, dtype=np.float64)
        ymode = {1: 0, 2: 1, 3: 2}[yref.ndim]
        p = default_params() if p is None else np.ascontiguousarray(p, dtype=np.float64)
        pmode = {1: 0, 2: 1, 3: 2}[p.ndim]
        tol = np.ascontiguousarray(np.broadcast_to(np.asarray(tol, dtype=np.float64), (4,)))
        sqp_it = np.zeros(B, dtype=np.int32)
        qp_it = np.zeros(B, dtype=np.int32)
        res = np.zeros((B, 4))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        rc = lib().orc_sqp_solve_batch(C.byref(self.o), _dp(self.X), _dp(self.U), _dp(x0), _dp(yref), ymode, _dp(p), pmode, int(max_iter),
                                       _dp(tol), ip(self.status), ip(sqp_it), ip(qp_it), _dp(res), B, self.nthreads)
        assert rc == 0
        self.iters = qp_it
        return self.U[:, 0].copy(), self.X.copy(), self.U.copy(), self.status.copy(), sqp_it, qp_it, res

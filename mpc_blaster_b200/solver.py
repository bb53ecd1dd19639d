"""Host-side mirror of the reference controller for the batched solve.

``BlasterMPC`` takes the same positional arguments as the reference's
``blasterModel(mass, J, l_x, l_y, N, Tf, c, Q, R, Q_t, blastThruster, statesBound,
controlBound)`` (reference src/scripts/blastermodel.py:16) and exposes the per-step
solve of src/scripts/simulation_blaster.py:56-105 as one batched call on torch CUDA
tensors.  All arithmetic happens in libmpcb.so (hand-written CUDA, include/mpcb.h); torch
only owns the device buffers and streams.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import MPCB_PER_INSTANCE, MPCB_PER_STAGE, MPCB_SHARED, MpcbConfig


def _diag(M, n, name):
    M = np.asarray(M, dtype=np.float64)
    if M.ndim == 1:
        d = M
    else:
        d = np.diag(M)
        if np.abs(M - np.diag(d)).max() != 0.0:
            raise ValueError(f"{name} must be diagonal (the reference's weights are; acados_ocp_blasterModel.json cost.W)")
    if d.shape[0] < n:
        raise ValueError(f"{name} has {d.shape[0]} entries, need {n}")
    return d[:n]


class MpcbError(RuntimeError):
    pass


def acados_json_args(path, N=None) -> dict:
    """Read the OCP data of an acados dump (the on-disk format the reference writes,
    blastermodel.py:289) into ``BlasterMPC`` constructor arguments.  dt = tf / dims.N is kept
    when the horizon is overridden."""
    import json
    d = json.load(open(path))
    nx, nu = int(d["dims"]["nx"]), int(d["dims"]["nu"])
    if (nx, nu) != (17, 6):
        raise ValueError("the dump is not the BLASTER OCP (nx, nu) = (17, 6)")
    so = d["solver_options"]
    if so["nlp_solver_type"] != "SQP_RTI" or so["integrator_type"] != "ERK" or so["hessian_approx"] != "GAUSS_NEWTON":
        raise ValueError("only the reference's SQP_RTI + ERK + GAUSS_NEWTON configuration is supported")
    if d["cost"]["cost_type"] != "LINEAR_LS" or d["cost"]["cost_type_e"] != "LINEAR_LS":
        raise ValueError("only LINEAR_LS costs are supported")
    W, We = np.array(d["cost"]["W"], dtype=np.float64), np.array(d["cost"]["W_e"], dtype=np.float64)
    cons = d["constraints"]
    if list(cons["idxbx"]) != list(range(nx)) or list(cons["idxbu"]) != list(range(nu)):
        raise ValueError("expected box bounds on every state and input")
    Nf = int(d["dims"]["N"])
    dt = float(so["tf"]) / Nf
    N = Nf if N is None else int(N)
    return dict(N=N, Tf=dt * N, Q=W[:nx, :nx], R=W[nx:, nx:], Q_t=We, blastThruster=float(d["parameter_values"][-1]),
                statesBound=np.array([cons["lbx"], cons["ubx"]]), controlBound=np.array([cons["lbu"], cons["ubu"]]),
                ipm_max_iter=int(so["qp_solver_iter_max"]))


def write_acados_json(path, *, N, Tf, Q, R, Q_t, blastThruster, statesBound, controlBound, ipm_max_iter=500, parameter_values=None):
    """Write OCP data in the layout of the acados dump the reference commits
    (src/scripts/acados_ocp_blasterModel.json, written at blastermodel.py:289): dims, LINEAR_LS
    cost (W, W_e, Vx, Vu, Vx_e as blastermodel.py:244-257 builds them), box bounds with their index
    sets (:261-270), default parameters (:280-282) and the solver options this solver implements
    (:272-287).  ``acados_json_args`` / ``BlasterMPC.from_acados_json`` read it back."""
    import json
    nx, nu = 17, 6
    Qd, Rd, Qtd = _diag(Q, nx, "Q"), _diag(R, nu, "R"), _diag(Q_t, nx, "Q_t")
    sb, cb = np.asarray(statesBound, dtype=np.float64), np.asarray(controlBound, dtype=np.float64)
    N, dt = int(N), float(Tf) / int(N)
    W = np.diag(np.concatenate([Qd, Rd]))
    Vx = np.vstack([np.eye(nx), np.zeros((nu, nx))])
    Vu = np.vstack([np.zeros((nx, nu)), np.eye(nu)])
    if parameter_values is None:
        pv = np.zeros(25)
        pv[24] = float(blastThruster)
    else:
        pv = np.asarray(parameter_values, dtype=np.float64).reshape(25)
    d = {"dims": {"N": N, "nx": nx, "nu": nu, "np": 25, "ny": nx + nu, "ny_e": nx, "nbx": nx, "nbu": nu, "nbx_0": nx, "nbx_e": 0},
         "cost": {"cost_type": "LINEAR_LS", "cost_type_0": "LINEAR_LS", "cost_type_e": "LINEAR_LS", "W": W.tolist(), "W_0": W.tolist(),
                  "W_e": np.diag(Qtd).tolist(), "Vx": Vx.tolist(), "Vu": Vu.tolist(), "Vx_e": np.eye(nx).tolist(),
                  "yref": [0.0] * (nx + nu), "yref_e": [0.0] * nx},
         "constraints": {"constr_type": "BGH", "lbx": sb[0].tolist(), "ubx": sb[1].tolist(), "lbu": cb[0].tolist(), "ubu": cb[1].tolist(),
                         "idxbx": list(range(nx)), "idxbu": list(range(nu)), "idxbxe_0": list(range(nx)), "idxbx_0": list(range(nx)),
                         "lbx_0": [0.0] * nx, "ubx_0": [0.0] * nx},
         "parameter_values": pv.tolist(),
         "solver_options": {"tf": dt * N, "time_steps": [dt] * N, "nlp_solver_type": "SQP_RTI", "integrator_type": "ERK",
                            "sim_method_num_stages": [4] * N, "sim_method_num_steps": [1] * N, "hessian_approx": "GAUSS_NEWTON",
                            "qp_solver": "PARTIAL_CONDENSING_HPIPM", "qp_solver_cond_N": N, "qp_solver_iter_max": int(ipm_max_iter),
                            "qp_solver_warm_start": 0, "globalization": "FIXED_STEP", "nlp_solver_step_length": 1.0,
                            "levenberg_marquardt": 0.0}}
    with open(path, "w") as f:
        json.dump(d, f, indent=1)
    return path


class BlasterMPC:
    """B independent BLASTER controllers solved together on one GPU.

    ``variant`` 17 = the reference's 17-state / 6-input model; 12 = QUAD12 (states 0..11,
    inputs 0..3, gimbal frozen); 13 = QUAT13 (QUAD12 with the attitude as a unit quaternion,
    x = [p, q(w,x,y,z), v, omega], quaternion algebra of the reference's utils/MathUtils.py).  ``blastThruster`` is stored but, exactly as in the
    reference (blastermodel.py:43), never used: the blast thrust is parameter p[24].
    """

    def __init__(self, mass, J, l_x, l_y, N, Tf, c, Q, R, Q_t, blastThruster, statesBound, controlBound, *,
                 batch: int = 1, variant: int = 17, dtype=torch.float64, device=None, ws_batch: int = 0, ipm_max_iter: int | None = None,
                 ipm_mu0: float = 1e2, ipm_thr0: float = -0.5, tol_stat: float = 1e-6, tol_eq: float = 1e-8,
                 tol_ineq: float = 1e-8, tol_comp: float = 1e-8, alpha_min: float = 1e-8, strict_reference: bool = False,
                 throughput_batch: int = 0, qp8_batch: int = 0, qp8_warps: int = 0):
        """``strict_reference`` restores the reference stack's solver semantics (mpcb_config.strict_reference,
        include/mpcb.h): explicit residual norms in the stopping test, no early exit on diverging multipliers, the last
        iterate applied on max-iter, and ``ipm_max_iter`` = 500 (blastermodel.py:279) unless given.  The default is
        the throughput-oriented rule set with a cap of 60 iterations."""
        if dtype != torch.float64:
            raise NotImplementedError("only float64 is implemented: the reference computes in IEEE double and an interior "
                                      "point with active state bounds is not viable in FP32 (DESIGN.md)")
        if not torch.cuda.is_available():
            raise MpcbError("BlasterMPC needs a CUDA device; this package has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise MpcbError("BlasterMPC needs a CUDA device; this package has no CPU fallback")
        if self.device.index is None:  # device='cuda' means the current CUDA device
            self.device = torch.device("cuda", torch.cuda.current_device())
        if ipm_max_iter is None:
            ipm_max_iter = 500 if strict_reference else 60
        if variant not in (17, 12, 13):
            raise ValueError("variant must be 17 (BLASTER17), 12 (QUAD12) or 13 (QUAT13)")
        self.nx, self.nu = {17: (17, 6), 12: (12, 4), 13: (13, 4)}[int(variant)]
        self.ny, self.N, self.batch = self.nx + self.nu, int(N), int(batch)
        self.blastThruster, self.variant = float(blastThruster), int(variant)
        cfg = MpcbConfig()
        cfg.variant, cfg.N, cfg.dt, cfg.mass = variant, int(N), float(Tf) / int(N), float(mass)
        cfg.J[:] = np.asarray(J, dtype=np.float64).reshape(9)
        cfg.l_x, cfg.l_y, cfg.c = float(l_x), float(l_y), float(c)
        sb = np.asarray(statesBound, dtype=np.float64)
        cb = np.asarray(controlBound, dtype=np.float64)
        for name, vals, n in (("Q", _diag(Q, self.nx, "Q"), self.nx), ("R", _diag(R, self.nu, "R"), self.nu),
                              ("Qt", _diag(Q_t, self.nx, "Q_t"), self.nx), ("lbx", sb[0], self.nx), ("ubx", sb[1], self.nx),
                              ("lbu", cb[0], self.nu), ("ubu", cb[1], self.nu)):
            arr = getattr(cfg, name)
            for i in range(n):
                arr[i] = float(vals[i])
        cfg.ipm_max_iter, cfg.ipm_mu0, cfg.ipm_thr0 = ipm_max_iter, ipm_mu0, ipm_thr0
        cfg.tol_stat, cfg.tol_eq, cfg.tol_ineq, cfg.tol_comp, cfg.alpha_min = tol_stat, tol_eq, tol_ineq, tol_comp, alpha_min
        cfg.dtype, cfg.max_batch, cfg.ws_batch, cfg.device = 64, self.batch, ws_batch, self.device.index
        cfg.strict_reference, cfg.throughput_batch, cfg.qp8_batch, cfg.qp8_warps = int(bool(strict_reference)), int(throughput_batch), int(qp8_batch), int(qp8_warps)
        self.strict_reference = bool(strict_reference)
        self.cfg = cfg
        self._h = C.c_void_p()
        if self.lib.mpcb_create(C.byref(cfg), C.byref(self._h)) != 0:
            raise MpcbError("mpcb_create: " + self.lib.mpcb_last_error(None).decode())
        self._yref = None

    # ------------------------------------------------------------------ helpers
    @classmethod
    def canonical(cls, N: int = 20, batch: int = 1, variant: int = 17, statesBound=None, controlBound=None, **kw):
        """Constants of reference simulation_blaster.py:12-30 with dt = 1/30 s (bounds can be overridden)."""
        Q = np.diag([1e3] * 6 + [5.0] * 3 + [10.0] * 3 + [1e-2] * 2 + [1e3] * 3)
        R = np.diag([5e-2] * 4 + [1e-5] * 2)
        sb = np.array([[-1.5, -1.5, 0, -0.174532925, -0.174532925, -0.349066, -1.0, -1.0, -1.0, -0.0872665, -0.0872665,
                        -0.0872665, -0.174532925, -0.523599, -1.5, -1.5, -2.5],
                       [1.5, 1.5, 5.0, 0.174532925, 0.174532925, 0.349066, 1.0, 1.0, 1.0, 0.0872665, 0.0872665, 0.0872665,
                        1.22173, 0.523599, 1.5, 1.5, 2.5]])
        cb = np.array([[0, 0, 0, 0, -0.0872665, -0.0872665], [65, 65, 65, 65, 0.0872665, 0.0872665]], dtype=np.float64)
        J = np.diag([0.50781, 0.47314, 0.72975])
        if variant == 13:
            # QUAT13: Euler weights -> quaternion components; Euler boxes (10, 10, 20 deg) -> boxes on the vector
            # part (sine of half the angle), q_w near 1; no gimbal, no POC states
            hq = np.sin(np.array([0.174532925, 0.174532925, 0.349066]) / 2)
            q = np.diag(Q)
            Q = np.diag(np.concatenate([q[0:3], [1e3] * 4, q[6:12]]))
            R = R[:4, :4]
            sb = np.array([np.concatenate([sb[0, 0:3], [0.9], -hq, sb[0, 6:12]]),
                           np.concatenate([sb[1, 0:3], [1.05], hq, sb[1, 6:12]])])
            cb = cb[:, :4]
        sb = sb if statesBound is None else np.asarray(statesBound, dtype=np.float64)
        cb = cb if controlBound is None else np.asarray(controlBound, dtype=np.float64)
        return cls(9.0, J, 0.3434, 0.3475, N, N / 30.0, 0.03, Q, R, 10 * Q, 2.2 * 9.81, sb, cb, batch=batch,
                   variant=variant, **kw)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self.lib.mpcb_destroy(h)
            self._h = None

    def _check(self, rc, what):
        if rc != 0:
            raise MpcbError(f"{what}: " + self.lib.mpcb_last_error(self._h).decode())

    def _t(self, a, shape_tail, name):
        """float64 contiguous CUDA tensor on this solver's device."""
        t = torch.as_tensor(a, dtype=torch.float64, device=self.device).contiguous()
        if tuple(t.shape[-len(shape_tail):]) != tuple(shape_tail):
            raise ValueError(f"{name}: expected trailing shape {shape_tail}, got {tuple(t.shape)}")
        return t

    def _mode(self, t, B, per_stage_rows, width, name):
        """SHARED / PER_INSTANCE / PER_STAGE from the exact shape of a torch tensor or NumPy array."""
        if t.ndim == 1 and tuple(t.shape) == (width,):
            return MPCB_SHARED
        if t.ndim == 2 and tuple(t.shape) == (B, width):
            return MPCB_PER_INSTANCE
        if t.ndim == 3 and tuple(t.shape) == (B, per_stage_rows, width):
            return MPCB_PER_STAGE
        raise ValueError(f"{name}: shape {tuple(t.shape)} is none of [{width}], [B,{width}], [B,{per_stage_rows},{width}]")

    @staticmethod
    def _stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    @staticmethod
    def _p(t):
        return C.c_void_p(0 if t is None else t.data_ptr())

    # ------------------------------------------------------------------ API
    def reset(self, x_init=None, u_init=None, B: int | None = None):
        """Set the stored SQP iterate (zeros by default, acados' initial iterate);
        x_init[B,nx] is copied to every stage, u_init[nu] or [B,nu] likewise."""
        B = self.batch if B is None else B
        with torch.cuda.device(self.device):
            xi = None if x_init is None else self._t(x_init, (self.nx,), "x_init").reshape(B, self.nx)
            ui = None if u_init is None else self._t(u_init, (self.nu,), "u_init")
            per = int(ui is not None and ui.dim() == 2)
            self._check(self.lib.mpcb_reset(self._h, self._p(xi), self._p(ui), per, B, self._stream()), "mpcb_reset")

    def shift(self, B: int | None = None):
        """Shift the stored iterate one stage forward (opt-in; the reference never shifts)."""
        B = self.batch if B is None else B
        with torch.cuda.device(self.device):
            self._check(self.lib.mpcb_shift(self._h, B, self._stream()), "mpcb_shift")

    def solve(self, x0, yref, p=None, want_traj: bool = True, sqp_iters: int = 1, sqp_tol=None):
        """One SQP-RTI iteration for every instance: returns (u0[B,nu], X[B,N+1,nx],
        U[B,N,nu], status[B] int32).  yref: [ny] | [B,ny] | [B,N+1,ny]; p: None | [25] |
        [B,25] | [B,N,25].  The iterate is kept in the handle for the next call (un-shifted
        warm start, as the reference's loop does).

        ``sqp_iters`` > 1 or ``sqp_tol`` given: multi-iteration SQP (mpcb_solve_sqp).  ``sqp_tol`` = a float or
        (stat, eq, ineq, comp): iterate to convergence, at most ``sqp_iters`` QPs per instance (the reference's
        options: nlp_solver_tol_* = 1e-6, nlp_solver_max_iter = 100); then ``status`` is the SQP status, ``self.iters``
        the summed interior-point iterations, ``self.sqp_iters`` the QPs solved and ``self.nlp_res[B,4]`` the residuals
        of the final iterate."""
        with torch.cuda.device(self.device):
            x0 = self._t(x0, (self.nx,), "x0")
            B = x0.shape[0] if x0.dim() == 2 else 1
            x0 = x0.reshape(B, self.nx)
            yref = self._t(yref, (self.ny,), "yref")
            ymode = self._mode(yref, B, self.N + 1, self.ny, "yref")
            pmode, pt = MPCB_SHARED, None
            if p is not None:
                pt = self._t(p, (25,), "p")
                pmode = self._mode(pt, B, self.N, 25, "p")
            u0 = torch.empty((B, self.nu), dtype=torch.float64, device=self.device)
            X = torch.empty((B, self.N + 1, self.nx), dtype=torch.float64, device=self.device) if want_traj else None
            U = torch.empty((B, self.N, self.nu), dtype=torch.float64, device=self.device) if want_traj else None
            status = torch.empty((B,), dtype=torch.int32, device=self.device)
            self.iters = torch.empty((B,), dtype=torch.int32, device=self.device)
            if sqp_iters == 1 and sqp_tol is None:
                rc = self.lib.mpcb_solve(self._h, self._p(x0), self._p(yref), ymode, self._p(pt), pmode, self._p(u0),
                                         self._p(X), self._p(U), self._p(status), self._p(self.iters), B, self._stream())
            else:  # SQP: re-linearise on the same data, to convergence when tolerances are given
                tol = None
                if sqp_tol is not None:
                    t = np.broadcast_to(np.asarray(sqp_tol, dtype=np.float64), (4,)).copy()
                    tol = (C.c_double * 4)(*t)
                self.sqp_iters = torch.empty((B,), dtype=torch.int32, device=self.device)
                self.nlp_res = torch.empty((B, 4), dtype=torch.float64, device=self.device)
                rc = self.lib.mpcb_solve_sqp(self._h, self._p(x0), self._p(yref), ymode, self._p(pt), pmode, int(sqp_iters),
                                             None if tol is None else C.cast(tol, C.c_void_p), self._p(u0), self._p(X), self._p(U),
                                             self._p(status), self._p(self.iters), self._p(self.sqp_iters), self._p(self.nlp_res),
                                             B, self._stream())
            self._check(rc, "mpcb_solve")
            self._yref = yref
        return u0, X, U, status

    @classmethod
    def from_acados_json(cls, path, *, N=None, batch=1, variant=17, mass=9.0, J=None, l_x=0.3434, l_y=0.3475, c=0.03, **kw):
        """Build a solver from an acados OCP dump such as the reference's
        src/scripts/acados_ocp_blasterModel.json (written at blastermodel.py:289): weights,
        bounds, horizon and tf are read from the file; the physical constants, which the dump does
        not contain (they are baked into the generated model code), come from the arguments
        (defaults = simulation_blaster.py:12-21)."""
        a = acados_json_args(path, N=N)
        J = np.diag([0.50781, 0.47314, 0.72975]) if J is None else J
        kw.setdefault("ipm_max_iter", a["ipm_max_iter"])  # qp_solver_iter_max of the dump (500 in the reference's) unless overridden
        return cls(mass, J, l_x, l_y, a["N"], a["Tf"], c, a["Q"], a["R"], a["Q_t"], a["blastThruster"], a["statesBound"],
                   a["controlBound"], batch=batch, variant=variant, **kw)

    def to_acados_json(self, path, parameter_values=None):
        """Write this controller's OCP data in the acados dump layout (see ``write_acados_json``)."""
        if self.variant != 17:
            raise ValueError("the acados dump format describes the 17-state / 6-input OCP")
        c, nx, nu = self.cfg, self.nx, self.nu
        return write_acados_json(path, N=self.N, Tf=float(c.dt) * self.N, Q=[c.Q[i] for i in range(nx)], R=[c.R[i] for i in range(nu)],
                                 Q_t=[c.Qt[i] for i in range(nx)], blastThruster=self.blastThruster,
                                 statesBound=[[c.lbx[i] for i in range(nx)], [c.ubx[i] for i in range(nx)]],
                                 controlBound=[[c.lbu[i] for i in range(nu)], [c.ubu[i] for i in range(nu)]],
                                 ipm_max_iter=int(c.ipm_max_iter), parameter_values=parameter_values)

    def solve_host(self, x0, yref, p=None, want_traj: bool = False):
        """Same through ``mpcb_solve_host``: NumPy in, NumPy out, copies inside the call."""
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        if x0.shape[-1] != self.nx or x0.ndim > 2:
            raise ValueError(f"x0: expected [B,{self.nx}] or [{self.nx}], got {x0.shape}")
        x0 = x0.reshape(-1, self.nx)
        B = x0.shape[0]
        # the library copies B*ny / B*(N+1)*ny / B*N*25 doubles straight from these pointers: validate the shapes first
        yref = np.ascontiguousarray(yref, dtype=np.float64)
        ymode = self._mode(yref, B, self.N + 1, self.ny, "yref")
        pmode, pp = MPCB_SHARED, None
        if p is not None:
            p = np.ascontiguousarray(p, dtype=np.float64)
            pmode, pp = self._mode(p, B, self.N, 25, "p"), p.ctypes.data
        u0 = np.empty((B, self.nu))
        X = np.empty((B, self.N + 1, self.nx)) if want_traj else None
        U = np.empty((B, self.N, self.nu)) if want_traj else None
        status = np.empty(B, dtype=np.int32)
        iters = np.empty(B, dtype=np.int32)
        self._check(self.lib.mpcb_solve_host(self._h, x0.ctypes.data, yref.ctypes.data, ymode, pp, pmode, u0.ctypes.data,
                                             None if X is None else X.ctypes.data, None if U is None else U.ctypes.data,
                                             status.ctypes.data, iters.ctypes.data, B), "mpcb_solve_host")
        self.iters_host = iters
        return u0, X, U, status

    def step_plant(self, x, u, p=None):
        """x+ = RK4(x, u, p) over dt (the reference's AcadosSimSolver)."""
        with torch.cuda.device(self.device):
            x = self._t(x, (self.nx,), "x")
            B = x.shape[0] if x.dim() == 2 else 1
            x = x.reshape(B, self.nx)
            u = self._t(u, (self.nu,), "u").reshape(B, self.nu)
            pmode, pt = MPCB_SHARED, None
            if p is not None:
                pt = self._t(p, (25,), "p")
                pmode = MPCB_SHARED if pt.dim() == 1 else MPCB_PER_INSTANCE
            xn = torch.empty_like(x)
            self._check(self.lib.mpcb_plant_step(self._h, self._p(x), self._p(u), self._p(pt), pmode, self._p(xn), B,
                                                 self._stream()), "mpcb_plant_step")
        return xn

    def closed_loop(self, x0, yref, p=None, steps: int = 1):
        """``steps`` control steps on the device: returns (x_final[B,nx], u_last[B,nu],
        n_fail[B], iters_sum[B])."""
        with torch.cuda.device(self.device):
            x = self._t(x0, (self.nx,), "x0").reshape(-1, self.nx).clone()
            B = x.shape[0]
            yref = self._t(yref, (self.ny,), "yref")
            ymode = self._mode(yref, B, self.N + 1, self.ny, "yref")
            pmode, pt = MPCB_SHARED, None
            if p is not None:
                pt = self._t(p, (25,), "p")
                pmode = MPCB_SHARED if pt.dim() == 1 else MPCB_PER_INSTANCE
            u_last = torch.empty((B, self.nu), dtype=torch.float64, device=self.device)
            n_fail = torch.empty((B,), dtype=torch.int32, device=self.device)
            iters = torch.empty((B,), dtype=torch.int32, device=self.device)
            self._check(self.lib.mpcb_closed_loop(self._h, self._p(x), self._p(yref), ymode, self._p(pt), pmode, int(steps),
                                                  self._p(u_last), self._p(n_fail), self._p(iters), B, self._stream()),
                        "mpcb_closed_loop")
            self._yref = yref
        return x, u_last, n_fail, iters

    def cost(self, yref=None, B: int | None = None):
        """Objective at the stored iterate (acados ``get_cost()``)."""
        with torch.cuda.device(self.device):
            yref = self._yref if yref is None else self._t(yref, (self.ny,), "yref")
            B = self.batch if B is None else B
            ymode = self._mode(yref, B, self.N + 1, self.ny, "yref")
            out = torch.empty((B,), dtype=torch.float64, device=self.device)
            self._check(self.lib.mpcb_cost(self._h, self._p(yref), ymode, self._p(out), B, self._stream()), "mpcb_cost")
        return out

    def iterate(self, B: int | None = None):
        B = self.batch if B is None else B
        with torch.cuda.device(self.device):
            X = torch.empty((B, self.N + 1, self.nx), dtype=torch.float64, device=self.device)
            U = torch.empty((B, self.N, self.nu), dtype=torch.float64, device=self.device)
            self._check(self.lib.mpcb_get_iterate(self._h, self._p(X), self._p(U), B, self._stream()), "mpcb_get_iterate")
        return X, U

    def set_iterate(self, X=None, U=None):
        with torch.cuda.device(self.device):
            Xt = None if X is None else self._t(X, (self.N + 1, self.nx), "X")
            Ut = None if U is None else self._t(U, (self.N, self.nu), "U")
            B = (Xt if Xt is not None else Ut).shape[0]
            self._check(self.lib.mpcb_set_iterate(self._h, self._p(Xt), self._p(Ut), B, self._stream()), "mpcb_set_iterate")

    def linearize(self, p=None, B: int | None = None):
        """Test hook: (A[B,N,nx,nx], B[B,N,nx,nu], b[B,N,nx]) of the stored iterate."""
        B = self.batch if B is None else B
        nz = self.nx + self.nu
        with torch.cuda.device(self.device):
            pmode, pt = MPCB_SHARED, None
            if p is not None:
                pt = self._t(p, (25,), "p")
                pmode = self._mode(pt, B, self.N, 25, "p")
            BAt = torch.empty((B, self.N, nz, self.nx), dtype=torch.float64, device=self.device)
            b = torch.empty((B, self.N, self.nx), dtype=torch.float64, device=self.device)
            self._check(self.lib.mpcb_debug_linearize(self._h, self._p(pt), pmode, self._p(BAt), self._p(b), B,
                                                      self._stream()), "mpcb_debug_linearize")
        return BAt[:, :, self.nu:, :].transpose(2, 3), BAt[:, :, :self.nu, :].transpose(2, 3), b

    def command_map(self, x, u0):
        """Attitude quaternion [w,x,y,z] and normalised thrust set-point
        (reference mavros_blaster_sim.py:27-30,91-100)."""
        with torch.cuda.device(self.device):
            x = self._t(x, (self.nx,), "x").reshape(-1, self.nx)
            u0 = self._t(u0, (self.nu,), "u0").reshape(-1, self.nu)
            B = x.shape[0]
            quat = torch.empty((B, 4), dtype=torch.float64, device=self.device)
            thrust = torch.empty((B,), dtype=torch.float64, device=self.device)
            self._check(self.lib.mpcb_command_map(self._h, self._p(x), self._p(u0), self._p(quat), self._p(thrust), B,
                                                  self._stream()), "mpcb_command_map")
        return quat, thrust

    def debug_qp(self, B: int | None = None):
        """Test hook (mpcb_debug_qp): the interior-point iterate the last ``solve`` ended with and the QP it solved, as a
        dict of tensors -- z, tl, tu, ll, lu, lb, ub, g [B,N+1,nz]; pi [B,N+1,nx]; BAt [B,N,nz,nx]; b [B,N,nx]."""
        B = self.batch if B is None else B
        nz, N, nx = self.nx + self.nu, self.N, self.nx
        with torch.cuda.device(self.device):
            mk = lambda *shape: torch.empty(shape, dtype=torch.float64, device=self.device)
            o = {k: mk(B, N + 1, nz) for k in ("z", "tl", "tu", "ll", "lu", "lb", "ub", "g")}
            o["pi"], o["BAt"], o["b"] = mk(B, N + 1, nx), mk(B, N, nz, nx), mk(B, N, nx)
            self._check(self.lib.mpcb_debug_qp(self._h, *[self._p(o[k]) for k in ("z", "pi", "tl", "tu", "ll", "lu", "lb", "ub", "g", "BAt", "b")],
                                               B, self._stream()), "mpcb_debug_qp")
        return o

    def profile(self, enable: bool = True):
        self._check(self.lib.mpcb_profile(self._h, int(enable)), "mpcb_profile")

    def last_kernel_ms(self):
        """(rollout kernel ms, QP kernel ms) of the last solve; needs profile(True)."""
        a, b = C.c_float(0), C.c_float(0)
        self._check(self.lib.mpcb_last_kernel_ms(self._h, C.byref(a), C.byref(b)), "mpcb_last_kernel_ms")
        return a.value, b.value

    def fp64_peak_tflops(self) -> float:
        v = C.c_double(0)
        self._check(self.lib.mpcb_fp64_peak(self.device.index, C.byref(v)), "mpcb_fp64_peak")
        return v.value

    def kernel_launches(self) -> int:
        return int(self.lib.mpcb_kernel_launches())

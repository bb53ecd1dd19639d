"""acados-shaped shim so the reference's closed-loop script body runs unchanged.

The loop body of reference src/scripts/simulation_blaster.py:56-105 drives two acados
objects, ``AcadosOcpSolver`` (set / cost_set / solve / get / get_cost) and
``AcadosSimSolver`` (set / solve / get).  ``blasterModel`` below has the reference's
constructor and ``generateModel()`` / ``generateController()`` methods
(blastermodel.py:16,47,214) and returns shims with those methods, backed by the batched
CUDA solver with B = 1.  Only the call patterns the reference uses are implemented;
anything else raises, rather than silently doing something different.
"""
from __future__ import annotations

import numpy as np
import torch

from .solver import BlasterMPC


class OcpSolverShim:
    def __init__(self, mpc: BlasterMPC):
        self.mpc = mpc
        N, nx, nu = mpc.N, mpc.nx, mpc.nu
        self._x0 = np.zeros(nx)
        self._yref = np.zeros((N + 1, nx + nu))
        self._p = np.zeros((N, 25))
        self._p[:, 24] = 2.2 * 9.81  # blastermodel.py:280-282
        self._lbx0 = self._ubx0 = None
        self._X = np.zeros((N + 1, nx))
        self._U = np.zeros((N, nu))
        self.status = 0

    def set(self, stage, field, value):
        v = np.asarray(value, dtype=np.float64).reshape(-1)
        if field in ("lbx", "ubx"):
            if stage != 0:
                raise NotImplementedError("only the stage-0 state pin (simulation_blaster.py:60-61) is supported")
            if field == "lbx":
                self._lbx0 = v.copy()
            else:
                self._ubx0 = v.copy()
        elif field == "p":
            self._p[stage] = v
        elif field == "x":
            self._X[stage] = v
            self.mpc.set_iterate(X=self._X[None])
        elif field == "u":
            self._U[stage] = v
            self.mpc.set_iterate(U=self._U[None])
        else:
            raise NotImplementedError(f"set(..., {field!r}, ...)")

    def cost_set(self, stage, field, value):
        if field != "yref":
            raise NotImplementedError(f"cost_set(..., {field!r}, ...)")
        v = np.asarray(value, dtype=np.float64).reshape(-1)
        self._yref[stage, :v.size] = v  # terminal stage takes the nx state references

    def solve(self):
        if self._lbx0 is None or self._ubx0 is None or np.any(self._lbx0 != self._ubx0):
            raise NotImplementedError("stage 0 must be pinned with lbx == ubx (idxbxe_0 in the reference's OCP)")
        u0, X, U, status = self.mpc.solve(self._lbx0[None], self._yref[None], self._p[None])
        self._X, self._U = X[0].cpu().numpy(), U[0].cpu().numpy()
        self.status = int(status[0].item())
        return self.status

    def get(self, stage, field):
        if field == "u":
            return self._U[stage].copy()
        if field == "x":
            return self._X[stage].copy()
        raise NotImplementedError(f"get(..., {field!r})")

    def get_cost(self):
        return float(self.mpc.cost(torch.as_tensor(self._yref[None]), B=1)[0].item())


class SimSolverShim:
    def __init__(self, mpc: BlasterMPC):
        self.mpc = mpc
        self._x = np.zeros(mpc.nx)
        self._u = np.zeros(mpc.nu)
        self._p = np.zeros(25)
        self._p[24] = 2.2 * 9.81

    def set(self, field, value):
        v = np.asarray(value, dtype=np.float64).reshape(-1)
        if field == "x":
            self._x = v.copy()
        elif field == "u":
            self._u = v.copy()
        elif field == "p":
            self._p = v.copy()
        else:
            raise NotImplementedError(f"set({field!r}, ...)")

    def solve(self):
        self._x = self.mpc.step_plant(self._x[None], self._u[None], self._p)[0].cpu().numpy()
        return 0

    def get(self, field):
        if field != "x":
            raise NotImplementedError(f"get({field!r})")
        return self._x.copy()


class blasterModel:  # noqa: N801  (the reference's class name)
    """Same constructor and methods as reference blastermodel.py:14-292."""

    def __init__(self, mass, J, l_x, l_y, N, Tf, c, Q, R, Q_t, blastThruster, statesBound, controlBound, **solver_kw):
        self._args = (mass, J, l_x, l_y, N, Tf, c, Q, R, Q_t, blastThruster, statesBound, controlBound)
        self._kw = solver_kw
        self._mpc = None

    def generateModel(self):
        return 0

    def generateController(self):
        self._mpc = BlasterMPC(*self._args, batch=1, **self._kw)
        return SimSolverShim(self._mpc), OcpSolverShim(self._mpc)

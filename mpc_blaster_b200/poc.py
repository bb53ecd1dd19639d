"""Jet point-of-contact (POC) Jacobian generator on the GPU: host-side mirror of the reference's
``Jacobian_POC_Solver`` (src/scripts/Jacobian_POC_Solver.py:18-300), batched.

Same constructor and method names as the reference class, so the lines of
simulation_blaster.py:37-39 run with only the import changed::

    solver = JacobianPOCSolver(150, 1, 0.000015)
    solver.initialise()
    J_mot, J_eul, J_pos = solver.getJacobians()

plus the batched entry points the reference does not have: ``solve_batch`` (B poses at once) and
``params_from_states`` (x[B,17] -> p[B,25], packed like simulation_blaster.py:67), which lets a
Monte-Carlo batch refresh its POC parameters every control step on the device instead of freezing
the hover-pose Jacobians for the whole run.  Everything goes through the C ABI
(``mpcb_poc_jacobians``); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

MODES = {"reference": 0, "analytic": 1}


class JacobianPOCSolver:
    def __init__(self, streamVelocity, M_c, Ts, *, mode: str = "reference", device=None):
        """streamVelocity [m/s], M_c (drag), Ts (unused by the computation, kept for the
        reference's signature, Jacobian_POC_Solver.py:20).  mode: "reference" = the reference's
        algorithm step for step, "analytic" = exact root and implicit-function Jacobians."""
        if not torch.cuda.is_available():
            raise RuntimeError("JacobianPOCSolver needs a CUDA device; this package has no CPU fallback")
        if mode not in MODES:
            raise ValueError(f"mode must be one of {sorted(MODES)}")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._streamVelocity, self._M_c, self._Ts, self.mode = float(streamVelocity), float(M_c), float(Ts), mode
        self._euler_angles, self._motor_angles, self._positions = np.zeros(3), np.zeros(2), np.zeros(3)
        self._POC = np.zeros(3)
        self._J_pos, self._J_eul, self._J_mot = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((3, 2))

    # ------------------------------------------------------------------ batched
    def _t(self, a, width):
        t = torch.as_tensor(a, dtype=torch.float64, device=self.device).reshape(-1, width).contiguous()
        return t

    @staticmethod
    def _p(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _launch(self, euler, motor, position, x17, B, T_blast, want_params):
        dev, f64 = self.device, torch.float64
        out = dict(poc=torch.empty((B, 3), dtype=f64, device=dev), J_mot=torch.empty((B, 3, 2), dtype=f64, device=dev),
                   J_eul=torch.empty((B, 3, 3), dtype=f64, device=dev), J_pos=torch.empty((B, 3, 3), dtype=f64, device=dev),
                   t_flight=torch.empty((B,), dtype=f64, device=dev), status=torch.empty((B,), dtype=torch.int32, device=dev))
        out["p"] = torch.empty((B, 25), dtype=f64, device=dev) if want_params else None
        with torch.cuda.device(dev):
            rc = self.lib.mpcb_poc_jacobians(self._p(euler), self._p(motor), self._p(position), self._p(x17), B, self._streamVelocity,
                                             self._M_c, MODES[self.mode], float(T_blast), self._p(out["poc"]), self._p(out["J_mot"]),
                                             self._p(out["J_eul"]), self._p(out["J_pos"]), self._p(out["p"]), self._p(out["t_flight"]),
                                             self._p(out["status"]), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise RuntimeError("mpcb_poc_jacobians: " + self.lib.mpcb_last_error(None).decode())
        return out

    def solve_batch(self, euler_angles, motor_angles, position, T_blast: float | None = None):
        """euler[B,3], motor[B,2], position[B,3] -> dict of CUDA tensors: poc[B,3], J_mot[B,3,2],
        J_eul[B,3,3], J_pos[B,3,3], t_flight[B], status[B] (and p[B,25] when T_blast is given)."""
        e, m, p = self._t(euler_angles, 3), self._t(motor_angles, 2), self._t(position, 3)
        if not (e.shape[0] == m.shape[0] == p.shape[0]):
            raise ValueError("euler, motor and position must have the same batch size")
        return self._launch(e, m, p, None, e.shape[0], 0.0 if T_blast is None else T_blast, T_blast is not None)

    def params_from_states(self, x, T_blast: float):
        """x[B,17] state vectors -> p[B,25] = [vec(J_mot), vec(J_eul), vec(J_pos), T_blast] (column-major
        blocks, simulation_blaster.py:67), evaluated at each vehicle's own pose."""
        x = self._t(x, 17)
        return self._launch(None, None, None, x, x.shape[0], T_blast, True)["p"]

    # ------------------------------------------------------------------ the reference's surface (B = 1)
    def initialise(self):
        """Jacobian_POC_Solver.py:53-57."""
        self.setInitConditions([0, 0, 0], [0, 0], [0, 0, 2])
        self.solveJacobians([0, 0, 0], [0, 0], [0, 0, 4])

    def setInitConditions(self, euler_angles, motor_angles, position):
        self._euler_angles = np.array(euler_angles, dtype=np.float64)
        self._motor_angles = np.array(motor_angles, dtype=np.float64)
        self._positions = np.array(position, dtype=np.float64)

    def solveJacobians(self, euler_angles, motor_angles, position):
        self.setInitConditions(euler_angles, motor_angles, position)
        o = self.solve_batch(self._euler_angles, self._motor_angles, self._positions)
        if int(o["status"][0]) != 0:
            raise RuntimeError(f"time-of-flight search failed (status {int(o['status'][0])})")
        self._POC = o["poc"][0].cpu().numpy()
        self._J_mot, self._J_eul, self._J_pos = (o[k][0].cpu().numpy() for k in ("J_mot", "J_eul", "J_pos"))
        self._Ts = float(o["t_flight"][0])

    def getJacobians(self):
        return self._J_mot, self._J_eul, self._J_pos

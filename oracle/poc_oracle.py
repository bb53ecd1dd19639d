"""TEST INFRASTRUCTURE -- CPU restatement of the reference's jet point-of-contact (POC) Jacobian
generator (SURVEY 8f "next" row 2).  Only tests/ may import this module; the product never does.

What it follows (all paths under /root/reference/src/scripts/):
  * nozzle pose: htm.py:7-36 (``compute_T_b_s2``, ``compute_T_w_b``) and
    Jacobian_POC_Solver.py:153-175 (``setInitConditions``): p0 = translation of T_w_b T_b_s2,
    v0 = R [0, 0, -streamVelocity];
  * jet model: Jacobian_POC_Solver.py:59-99: p' = v, v' = -M_c v + g, explicit RK4, 10 steps over T;
  * time of flight: :115-152: Newton on z(T) = 0 from T = 0.1 with a forward-difference slope
    (dT = 1e-5), a negative iterate is reflected, stop at |z| <= 1e-3;
  * Jacobians: :234-300: forward differences, eps = 1e-6, over the Euler angles, the two nozzle
    angles and the position, each coordinate on its own (the reference's list call pattern, see
    tests/golden/make_poc_golden.py).

PARITY PINNED by tests/golden/poc_golden.npz, produced by running the reference's own
Jacobian_POC_Solver.py / htm.py (acados' ERK integrator replaced by a stub implementing the
scheme the reference configures).

``analytic_jacobians`` is the exact counterpart the CUDA path also offers: closed-form flight of
the linear jet ODE, Newton to machine precision, Jacobians by the implicit-function theorem.
"""
from __future__ import annotations

import numpy as np

G = 9.81
EPS_FD = 1e-6          # Jacobian_POC_Solver.py:37
ROOT_TOL = 1e-3        # :134
ROOT_DT = 1e-5         # :145
ROOT_T0 = 0.1          # :243
RK_STEPS = 10          # :95


def T_b_s2(a1: float, a2: float) -> np.ndarray:
    """htm.py:7-29: body -> swivel 1 -> swivel 2 -> nozzle."""
    hbs1 = np.array([[1, 0, 0, 0.01672], [0, 1, 0, 0], [0, 0, 1, -0.22937], [0, 0, 0, 1.0]])
    c1, s1, c2, s2 = np.cos(a1), np.sin(a1), np.cos(a2), np.sin(a2)
    hs1s2 = np.array([[c1, 0, s1, 0.0425], [0, 1, 0, 0], [-s1, 0, c1, 0], [0, 0, 0, 1.0]])
    hs2n = np.array([[1, 0, 0, -0.05322], [0, c2, s2, 0], [0, -s2, c2, -0.15946], [0, 0, 0, 1.0]])
    return hbs1 @ hs1s2 @ hs2n


def rot_w_b(phi: float, theta: float, psi: float) -> np.ndarray:
    """htm.py:31-36: scipy ``Rotation.from_euler('zyx', [psi, theta, phi])`` -- lower case = extrinsic
    rotations about the fixed z, then y, then x axis, i.e. R = Rx(phi) Ry(theta) Rz(psi).  (The OCP
    model uses Rz Ry Rx, blastermodel.py:103-122; the reference mixes the two conventions.)"""
    cf, sf, ct, st, cp, sp_ = np.cos(phi), np.sin(phi), np.cos(theta), np.sin(theta), np.cos(psi), np.sin(psi)
    Rx = np.array([[1, 0, 0], [0, cf, -sf], [0, sf, cf]])
    Ry = np.array([[ct, 0, st], [0, 1, 0], [-st, 0, ct]])
    Rz = np.array([[cp, -sp_, 0], [sp_, cp, 0], [0, 0, 1]])
    return Rx @ Ry @ Rz


def T_w_b(phi, theta, psi, position) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = rot_w_b(phi, theta, psi)
    T[:3, 3] = position
    return T


def init_conditions(euler, motor, position, stream_velocity: float) -> np.ndarray:
    """Jacobian_POC_Solver.py:153-175."""
    T = T_w_b(euler[0], euler[1], euler[2], position) @ T_b_s2(motor[0], motor[1])
    return np.hstack([T[:3, 3], T[:3, :3] @ np.array([0.0, 0.0, -stream_velocity])])


def flight_rk4(x0: np.ndarray, T: float, M_c: float) -> np.ndarray:
    """Jacobian_POC_Solver.py:77-99: ERK, 4 stages, 10 steps."""
    g = np.array([0.0, 0.0, -G])
    f = lambda x: np.hstack([x[3:], -M_c * x[3:] + g])
    h = T / RK_STEPS
    x = np.array(x0, dtype=np.float64)
    for _ in range(RK_STEPS):
        k1 = f(x); k2 = f(x + 0.5 * h * k1); k3 = f(x + 0.5 * h * k2); k4 = f(x + h * k3)
        x = x + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    return x


def time_of_flight(x0: np.ndarray, M_c: float, T0: float = ROOT_T0) -> float:
    """Jacobian_POC_Solver.py:115-152."""
    fun = lambda T: flight_rk4(x0, T, M_c)[2]
    T_N = T0
    while True:
        f = fun(T_N)
        fp = (fun(T_N + ROOT_DT) - f) / ROOT_DT
        T_N = T_N - f / fp
        if T_N < 0:
            T_N = -T_N
        if abs(fun(T_N)) <= ROOT_TOL:
            return T_N


def poc(euler, motor, position, stream_velocity: float, M_c: float) -> np.ndarray:
    x0 = init_conditions(euler, motor, position, stream_velocity)
    return flight_rk4(x0, time_of_flight(x0, M_c), M_c)[:3]


def solve_jacobians(euler, motor, position, stream_velocity: float = 150.0, M_c: float = 1.0):
    """Jacobian_POC_Solver.py:234-300 -> (POC[3], J_mot[3,2], J_eul[3,3], J_pos[3,3])."""
    euler, motor, position = (np.array(a, dtype=np.float64) for a in (euler, motor, position))
    base = poc(euler, motor, position, stream_velocity, M_c)
    J_eul, J_mot, J_pos = np.zeros((3, 3)), np.zeros((3, 2)), np.zeros((3, 3))
    for i in range(3):
        e = euler.copy(); e[i] = e[i] + EPS_FD
        J_eul[:, i] = (poc(e, motor, position, stream_velocity, M_c) - base) / EPS_FD
    for i in range(2):
        m = motor.copy(); m[i] = m[i] + EPS_FD
        J_mot[:, i] = (poc(euler, m, position, stream_velocity, M_c) - base) / EPS_FD
    for i in range(3):
        p = position.copy(); p[i] = p[i] + EPS_FD
        J_pos[:, i] = (poc(euler, motor, p, stream_velocity, M_c) - base) / EPS_FD
    return base, J_mot, J_eul, J_pos


# ------------------------------------------------------------------ exact counterpart
def flight_closed_form(x0: np.ndarray, T: float, M_c: float):
    """p(T), v(T) of p' = v, v' = -c v + g:  v = vinf + (v0 - vinf) e^{-cT}, p = p0 + vinf T + (v0 - vinf)(1 - e^{-cT})/c."""
    p0, v0 = x0[:3], x0[3:]
    vinf = np.array([0.0, 0.0, -G]) / M_c
    e = np.exp(-M_c * T)
    k = -np.expm1(-M_c * T) / M_c
    return p0 + vinf * T + (v0 - vinf) * k, vinf + (v0 - vinf) * e, k


def analytic_jacobians(euler, motor, position, stream_velocity: float = 150.0, M_c: float = 1.0):
    """Exact POC (z(T*) = 0 to machine precision) and d POC / d (nozzle angles, Euler angles, position)."""
    euler, motor, position = (np.array(a, dtype=np.float64) for a in (euler, motor, position))
    x0 = init_conditions(euler, motor, position, stream_velocity)
    T = max(x0[2] / max(-x0[5], 1e-9), 1e-6)
    for _ in range(50):
        p, v, _ = flight_closed_form(x0, T, M_c)
        dT = -p[2] / v[2]
        T += dT
        if abs(dT) <= 1e-15 * max(T, 1.0):
            break
    p, v, k = flight_closed_form(x0, T, M_c)

    def dx0(fun, h=1e-7):  # derivative of the initial conditions: central differences are exact enough for a checker
        return (fun(h) - fun(-h)) / (2 * h)
    cols = []
    for which, n in (("m", 2), ("e", 3), ("p", 3)):
        for i in range(n):
            def f(h, which=which, i=i):
                e_, m_, p_ = euler.copy(), motor.copy(), position.copy()
                {"e": e_, "m": m_, "p": p_}[which][i] += h
                return init_conditions(e_, m_, p_, stream_velocity)
            d = dx0(f)
            dp = d[:3] + d[3:] * k            # d p(T)/d theta at fixed T
            cols.append(dp - v * dp[2] / v[2])  # implicit-function theorem on z(T*(theta), theta) = 0
    J = np.stack(cols, axis=1)
    return p, J[:, 0:2], J[:, 2:5], J[:, 5:8], T

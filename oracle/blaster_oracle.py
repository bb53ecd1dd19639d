"""NumPy FP64 oracle of the BLASTER quadrotor SQP-RTI solve.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (solver semantics) -- see ``oracle/__init__.py``.

This is a CPU restatement of the reference's per-control-step optimal-control
solve.  The reference itself only *describes* the OCP and hands it to acados:

* dynamics            /root/reference/src/scripts/blastermodel.py:93-167,171-210
* cost (LINEAR_LS)    blastermodel.py:228-257
* bounds              blastermodel.py:261-270
* solver options      blastermodel.py:272-287 and
                      src/scripts/acados_ocp_blasterModel.json (solver_options)
* closed-loop driver  src/scripts/simulation_blaster.py:56-105

acados / HPIPM conventions that are *not* citeable inside /root/reference
 are marked ``[upstream Dn]`` where they are used.

Deliberately simple and structurally different from the product: the QP is
solved with a *dense-KKT* Mehrotra interior point (the CUDA path and the C
oracle use a stage-wise Riccati factorisation), so that agreement between the
two is evidence, not tautology.  The QP is strictly convex, so its primal
solution is unique and algorithm-independent.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

GRAVITY = 9.81  # blastermodel.py:93  gravity = [0, 0, -9.81] (ENU)


# --------------------------------------------------------------------------
# problem data
# --------------------------------------------------------------------------
@dataclass
class BlasterProblem:
    """POD mirror of the ``blasterModel`` constructor (blastermodel.py:16-45).

    ``variant`` 17 = the reference's model (17 states / 6 inputs / 25 params),
    12 = QUAD12: states 0..11 and inputs 0..3 of the same model with the gimbal
    frozen at alpha1 = alpha2 = 0;
    13 = QUAT13: QUAD12 with the attitude as a unit quaternion instead of Euler angles
    (exists nowhere in the reference -- see ``f13``).
    """

    mass: float
    J: np.ndarray  # 3x3
    l_x: float
    l_y: float
    c: float
    N: int
    dt: float
    Q: np.ndarray  # diag, nx
    R: np.ndarray  # diag, nu
    Qt: np.ndarray  # diag, nx
    lbx: np.ndarray
    ubx: np.ndarray
    lbu: np.ndarray
    ubu: np.ndarray
    variant: int = 17
    Jinv: np.ndarray = field(init=False)

    def __post_init__(self):
        self.J = np.asarray(self.J, dtype=np.float64).reshape(3, 3)
        self.Jinv = np.linalg.inv(self.J)
        for k in ("Q", "R", "Qt", "lbx", "ubx", "lbu", "ubu"):
            setattr(self, k, np.asarray(getattr(self, k), dtype=np.float64).copy())

    @property
    def nx(self):
        return {17: 17, 13: 13}.get(self.variant, 12)

    @property
    def nu(self):
        return 6 if self.variant == 17 else 4

    @property
    def np_(self):
        return 25

    @property
    def ny(self):
        return self.nx + self.nu


def canonical_problem(N: int = 20, variant: int = 17) -> BlasterProblem:
    """Constants of simulation_blaster.py:12-30 (identical to the committed JSON
    dump); dt = Tf/N = 2.0/60 = 1/30 s is kept fixed when N changes."""
    Q = np.array([1e3] * 6 + [5.0] * 3 + [10.0] * 3 + [1e-2] * 2 + [1e3] * 3)
    R = np.array([5e-2] * 4 + [1e-5] * 2)
    lbx = np.array([-1.5, -1.5, 0, -0.174532925, -0.174532925, -0.349066, -1.0, -1.0, -1.0,
                    -0.0872665, -0.0872665, -0.0872665, -0.174532925, -0.523599, -1.5, -1.5, -2.5])
    ubx = np.array([1.5, 1.5, 5.0, 0.174532925, 0.174532925, 0.349066, 1.0, 1.0, 1.0,
                    0.0872665, 0.0872665, 0.0872665, 1.22173, 0.523599, 1.5, 1.5, 2.5])
    lbu = np.array([0, 0, 0, 0, -0.0872665, -0.0872665], dtype=np.float64)
    ubu = np.array([65, 65, 65, 65, 0.0872665, 0.0872665], dtype=np.float64)
    if variant == 13:
        # QUAT13: x = [p, q(w,x,y,z), v, omega].  Weights of the Euler angles go to the quaternion
        # components; the Euler-angle boxes (10, 10, 20 deg) become boxes on the vector part
        # (sin of half the angle) and q_w stays near 1.
        Q13 = np.concatenate([Q[0:3], [1e3] * 4, Q[6:12]])
        lb13 = np.concatenate([lbx[0:3], [0.9, -np.sin(0.174532925 / 2), -np.sin(0.174532925 / 2), -np.sin(0.349066 / 2)], lbx[6:12]])
        ub13 = np.concatenate([ubx[0:3], [1.05, np.sin(0.174532925 / 2), np.sin(0.174532925 / 2), np.sin(0.349066 / 2)], ubx[6:12]])
        return BlasterProblem(
            mass=9.0, J=np.diag([0.50781, 0.47314, 0.72975]), l_x=0.3434, l_y=0.3475, c=0.03,
            N=N, dt=2.0 / 60, Q=Q13, R=R[:4], Qt=10 * Q13, lbx=lb13, ubx=ub13, lbu=lbu[:4], ubu=ubu[:4], variant=13)
    nx, nu = (17, 6) if variant == 17 else (12, 4)
    return BlasterProblem(
        mass=9.0, J=np.diag([0.50781, 0.47314, 0.72975]), l_x=0.3434, l_y=0.3475, c=0.03,
        N=N, dt=2.0 / 60, Q=Q[:nx], R=R[:nu], Qt=10 * Q[:nx],
        lbx=lbx[:nx], ubx=ubx[:nx], lbu=lbu[:nu], ubu=ubu[:nu], variant=variant)


def default_params() -> np.ndarray:
    """blastermodel.py:280-282: all POC Jacobians 0, T_blast = 2.2*9.81."""
    p = np.zeros(25)
    p[24] = 2.2 * 9.81
    return p


def canonical_x0_yref():
    """simulation_blaster.py:47-48."""
    x0 = np.zeros(17)
    yref = np.zeros(23)
    yref[2] = 3.5
    yref[14] = 0.2
    return x0, yref


def pack_params(J_mot, J_eul, J_pos, T_blast):
    """simulation_blaster.py:67: column-major vec of J_mot(3x2), J_eul(3x3),
    J_pos(3x3) then T_blast  (= blastermodel.py:203-210 with casadi reshape)."""
    return np.concatenate([np.asarray(J_mot).reshape(-1, order="F"),
                           np.asarray(J_eul).reshape(-1, order="F"),
                           np.asarray(J_pos).reshape(-1, order="F"), [T_blast]])


# --------------------------------------------------------------------------
# A1/A2: continuous dynamics and Jacobians (17-state model)
# --------------------------------------------------------------------------
def _moment_map(P: BlasterProblem) -> np.ndarray:
    """blastermodel.py:95-101: M_tau = G @ T."""
    ly, lx, c = P.l_y, P.l_x, P.c
    return np.array([[-ly, ly, -ly, ly],
                     [-lx, lx, lx, -lx],
                     [-c, -c, c, c]])


def _rot(phi, th, psi):
    """R = Rz(psi) Ry(theta) Rx(phi)  (blastermodel.py:103-122)."""
    cf, sf, ct, st, cp, sp = np.cos(phi), np.sin(phi), np.cos(th), np.sin(th), np.cos(psi), np.sin(psi)
    return np.array([[cp * ct, cp * st * sf - sp * cf, cp * st * cf + sp * sf],
                     [sp * ct, sp * st * sf + cp * cf, sp * st * cf - cp * sf],
                     [-st, ct * sf, ct * cf]])


def _euler_rate_map(phi, th):
    """inv(R_to_omega) of blastermodel.py:128-140,162 in closed form."""
    cf, sf, ct, tt = np.cos(phi), np.sin(phi), np.cos(th), np.tan(th)
    return np.array([[1.0, sf * tt, cf * tt],
                     [0.0, cf, -sf],
                     [0.0, sf / ct, cf / ct]])


def f17(x, u, p, P: BlasterProblem):
    """xdot = f(x,u,p), blastermodel.py:124,162-167,191-201."""
    phi, th, psi = x[3], x[4], x[5]
    v, om = x[6:9], x[9:12]
    a1, a2 = x[12], x[13]
    T, adot = u[0:4], u[4:6]
    Jang = p[0:6].reshape(3, 2, order="F")
    Jeul = p[6:15].reshape(3, 3, order="F")
    Jp = p[15:24].reshape(3, 3, order="F")
    Tb = p[24]
    Rm = _rot(phi, th, psi)
    g3 = np.array([np.sin(a1) * np.cos(a2), -np.sin(a2), np.cos(a1) * np.cos(a2)])  # R_gimbal e3
    w = np.array([0.0, 0.0, T.sum()]) + Tb * g3
    etad = _euler_rate_map(phi, th) @ om
    vd = (Rm @ w) / P.mass + np.array([0.0, 0.0, -GRAVITY])
    omd = P.Jinv @ (_moment_map(P) @ T - np.cross(om, P.J @ om))
    pocd = Jp @ v + Jeul @ etad + Jang @ adot
    return np.concatenate([v, etad, vd, omd, adot, pocd])


def _skew(a):
    return np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])


def jac17(x, u, p, P: BlasterProblem):
    """Analytic (df/dx, df/du) of ``f17`` (what acados obtains from CasADi's
    forward VDE because integrator_type='ERK', blastermodel.py:277)."""
    phi, th, psi = x[3], x[4], x[5]
    v, om = x[6:9], x[9:12]
    a1, a2 = x[12], x[13]
    T = u[0:4]
    Jang = p[0:6].reshape(3, 2, order="F")
    Jeul = p[6:15].reshape(3, 3, order="F")
    Jp = p[15:24].reshape(3, 3, order="F")
    Tb = p[24]
    cf, sf, ct, st, cp, sp = np.cos(phi), np.sin(phi), np.cos(th), np.sin(th), np.cos(psi), np.sin(psi)
    tt = st / ct
    Rm = _rot(phi, th, psi)
    g3 = np.array([np.sin(a1) * np.cos(a2), -np.sin(a2), np.cos(a1) * np.cos(a2)])
    w = np.array([0.0, 0.0, T.sum()]) + Tb * g3
    m = 1.0 / P.mass
    fx = np.zeros((17, 17))
    fu = np.zeros((17, 6))
    # pdot = v
    fx[0:3, 6:9] = np.eye(3)
    # euler rates
    E = _euler_rate_map(phi, th)
    etad = E @ om
    fd, td, pd = etad
    dE_dphi_om = np.array([tt * td, -pd * ct, td / ct])
    dE_dth_om = np.array([pd / ct, 0.0, pd * tt])
    fx[3:6, 3] = dE_dphi_om
    fx[3:6, 4] = dE_dth_om
    fx[3:6, 9:12] = E
    # vdot = m R w + g
    dR_dphi_w = Rm[:, 2] * w[1] - Rm[:, 1] * w[2]
    dR_dth = np.array([[-cp * st, cp * ct * sf, cp * ct * cf],
                       [-sp * st, sp * ct * sf, sp * ct * cf],
                       [-ct, -st * sf, -st * cf]])
    Rw = Rm @ w
    fx[6:9, 3] = m * dR_dphi_w
    fx[6:9, 4] = m * (dR_dth @ w)
    fx[6:9, 5] = m * np.array([-Rw[1], Rw[0], 0.0])
    dg_da1 = np.array([np.cos(a1) * np.cos(a2), 0.0, -np.sin(a1) * np.cos(a2)])
    dg_da2 = np.array([-np.sin(a1) * np.sin(a2), -np.cos(a2), -np.cos(a1) * np.sin(a2)])
    fx[6:9, 12] = m * Tb * (Rm @ dg_da1)
    fx[6:9, 13] = m * Tb * (Rm @ dg_da2)
    fu[6:9, 0:4] = m * Rm[:, 2][:, None]
    # omegadot = Jinv (G T - om x J om)
    fx[9:12, 9:12] = -P.Jinv @ (_skew(om) @ P.J - _skew(P.J @ om))
    fu[9:12, 0:4] = P.Jinv @ _moment_map(P)
    # alphadot = u[4:6]
    fu[12, 4] = 1.0
    fu[13, 5] = 1.0
    # pocdot = Jp v + Jeul etad + Jang adot
    fx[14:17, 6:9] = Jp
    fx[14:17, 3] = Jeul @ dE_dphi_om
    fx[14:17, 4] = Jeul @ dE_dth_om
    fx[14:17, 9:12] = Jeul @ E
    fu[14:17, 4:6] = Jang
    return fx, fu


# --------------------------------------------------------------------------
# QUAT13: the 12-state quadrotor with a quaternion attitude.
# Not in the reference: its utils/MathUtils.py provides the quaternion algebra
# (quatMultiplication :5-23, unitQuatInversion :25-39, quat2Rot :41-54) but no model uses
# it.  PARITY UNPINNED for the model itself; it is tied to the Euler model by
# construction (same forces and moments, R(q) = Rz Ry Rx at q = q(phi,theta,psi)) and
# tested against it.  x = [p(3), q(w,x,y,z), v(3), omega(3)], u = [T0..T3].
# --------------------------------------------------------------------------
def quat_mul(a, b):
    """MathUtils.quatMultiplication (Hamilton product, q = [w,x,y,z])."""
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                     a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                     a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def quat_to_rot(e):
    """MathUtils.quat2Rot."""
    w, x, y, z = e
    return np.array([[2 * (w * w + x * x) - 1, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 2 * (w * w + y * y) - 1, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 2 * (w * w + z * z) - 1]])


def euler_to_quat(phi, th, psi):
    """Quaternion of R = Rz(psi) Ry(theta) Rx(phi) (blastermodel.py:122)."""
    cf, sf, ct, st, cp, sp = np.cos(phi / 2), np.sin(phi / 2), np.cos(th / 2), np.sin(th / 2), np.cos(psi / 2), np.sin(psi / 2)
    return np.array([cp * ct * cf + sp * st * sf, cp * ct * sf - sp * st * cf, cp * st * cf + sp * ct * sf, sp * ct * cf - cp * st * sf])


def x12_to_x13(x12):
    """[p, euler, v, omega] -> [p, q, v, omega]."""
    x12 = np.asarray(x12, dtype=np.float64)
    return np.concatenate([x12[0:3], euler_to_quat(*x12[3:6]), x12[6:12]])


def f13(x, u, p, P: BlasterProblem):
    """pdot = v;  qdot = 1/2 q (x) [0, omega];  vdot = R(q) e3 (sum T + T_blast)/M + g  (the jet acts along
    body z: gimbal frozen at 0, as in QUAD12);  omegadot as blastermodel.py:164."""
    q, v, om, T = x[3:7], x[7:10], x[10:13], u[0:4]
    qd = 0.5 * quat_mul(q, np.array([0.0, om[0], om[1], om[2]]))
    vd = quat_to_rot(q)[:, 2] * (T.sum() + p[24]) / P.mass + np.array([0.0, 0.0, -GRAVITY])
    omd = P.Jinv @ (_moment_map(P) @ T - np.cross(om, P.J @ om))
    return np.concatenate([v, qd, vd, omd])


def jac13(x, u, p, P: BlasterProblem):
    """Analytic (df/dx, df/du) of ``f13``."""
    (w, qx, qy, qz), om, T = x[3:7], x[10:13], u[0:4]
    a, b, c = om
    F = (T.sum() + p[24]) / P.mass
    fx, fu = np.zeros((13, 13)), np.zeros((13, 4))
    fx[0:3, 7:10] = np.eye(3)
    fx[3:7, 3:7] = 0.5 * np.array([[0, -a, -b, -c], [a, 0, c, -b], [b, -c, 0, a], [c, b, -a, 0]])
    fx[3:7, 10:13] = 0.5 * np.array([[-qx, -qy, -qz], [w, -qz, qy], [qz, w, -qx], [-qy, qx, w]])
    fx[7:10, 3:7] = F * 2.0 * np.array([[qy, qz, w, qx], [-qx, -w, qz, qy], [2 * w, 0, 0, 2 * qz]])
    fu[7:10, 0:4] = (quat_to_rot(x[3:7])[:, 2] / P.mass)[:, None]
    fx[10:13, 10:13] = -P.Jinv @ (_skew(om) @ P.J - _skew(P.J @ om))
    fu[10:13, 0:4] = P.Jinv @ _moment_map(P)
    return fx, fu


def _pad(x, u, P):
    if P.variant == 17:
        return x, u
    xx = np.zeros(17)
    xx[:12] = x
    uu = np.zeros(6)
    uu[:4] = u
    return xx, uu


def f(x, u, p, P: BlasterProblem):
    if P.variant == 13:
        return f13(np.asarray(x, dtype=np.float64), np.asarray(u, dtype=np.float64), p, P)
    xx, uu = _pad(x, u, P)
    return f17(xx, uu, p, P)[:P.nx]


def jac(x, u, p, P: BlasterProblem):
    if P.variant == 13:
        return jac13(np.asarray(x, dtype=np.float64), np.asarray(u, dtype=np.float64), p, P)
    xx, uu = _pad(x, u, P)
    fx, fu = jac17(xx, uu, p, P)
    return fx[:P.nx, :P.nx], fu[:P.nx, :P.nu]


# --------------------------------------------------------------------------
# A3: ERK4 + forward sensitivities
# --------------------------------------------------------------------------
def rk4_sens(x, u, p, P: BlasterProblem):
    """One classic RK4 step of length dt of the augmented ODE [x; S]
    (Sdot = fx S + [0 fu], S(0) = [I 0])  [upstream D6]; ERK, 4 stages, 1 step
    per shooting interval (JSON sim_method_num_stages/num_steps).
    Returns x+, A = S[:, :nx], B = S[:, nx:]."""
    nx, nu, h = P.nx, P.nu, P.dt
    S0 = np.hstack([np.eye(nx), np.zeros((nx, nu))])

    def rhs(xs, Ss):
        fx, fu = jac(xs, u, p, P)
        dS = fx @ Ss
        dS[:, nx:] += fu
        return f(xs, u, p, P), dS

    k1, K1 = rhs(x, S0)
    k2, K2 = rhs(x + 0.5 * h * k1, S0 + 0.5 * h * K1)
    k3, K3 = rhs(x + 0.5 * h * k2, S0 + 0.5 * h * K2)
    k4, K4 = rhs(x + h * k3, S0 + h * K3)
    xn = x + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    S = S0 + h / 6.0 * (K1 + 2 * K2 + 2 * K3 + K4)
    return xn, S[:, :nx], S[:, nx:]


def plant_step(x, u, p, P: BlasterProblem):
    """AcadosSimSolver built from the same OCP (blastermodel.py:290): same RK4
    step, T = tf/N (simulation_blaster.py:94-104)."""
    h = P.dt
    k1 = f(x, u, p, P)
    k2 = f(x + 0.5 * h * k1, u, p, P)
    k3 = f(x + 0.5 * h * k2, u, p, P)
    k4 = f(x + h * k3, u, p, P)
    return x + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)


# --------------------------------------------------------------------------
# A4/A5: Gauss-Newton QP of one RTI iteration
# --------------------------------------------------------------------------
@dataclass
class StageQP:
    """min sum_k 1/2 dz_k' diag(Hk) dz_k + g_k' dz_k,  dz_k = [du_k; dx_k]
    s.t. dx_{k+1} = A_k dx_k + B_k du_k + b_k,  dx_0 given, box bounds on dz."""
    A: np.ndarray  # [N, nx, nx]
    B: np.ndarray  # [N, nx, nu]
    b: np.ndarray  # [N, nx]
    Hu: np.ndarray  # [N, nu] diag
    Hx: np.ndarray  # [N+1, nx] diag
    gu: np.ndarray  # [N, nu]
    gx: np.ndarray  # [N+1, nx]
    lbu: np.ndarray  # [N, nu]   bounds on du
    ubu: np.ndarray
    lbx: np.ndarray  # [N+1, nx] bounds on dx (+-inf where absent)
    ubx: np.ndarray
    dx0: np.ndarray  # [nx]


def expand_yref(yref, P: BlasterProblem):
    """Accept yref[ny] (same vector on every stage, as simulation_blaster.py:63-78
    does) or yref[N+1, ny] (per stage; the terminal row uses its first nx)."""
    yref = np.asarray(yref, dtype=np.float64)
    if yref.ndim == 1:
        yref = np.broadcast_to(yref, (P.N + 1, P.ny))
    assert yref.shape == (P.N + 1, P.ny)
    return yref


def expand_p(p, P: BlasterProblem):
    p = default_params() if p is None else np.asarray(p, dtype=np.float64)
    if p.ndim == 1:
        p = np.broadcast_to(p, (P.N, 25))
    assert p.shape == (P.N, 25)
    return p


def build_qp(X, U, x0, yref, p, P: BlasterProblem) -> StageQP:
    N, nx, nu, dt = P.N, P.nx, P.nu, P.dt
    yref = expand_yref(yref, P)
    p = expand_p(p, P)
    A = np.zeros((N, nx, nx))
    B = np.zeros((N, nx, nu))
    b = np.zeros((N, nx))
    for k in range(N):
        xn, A[k], B[k] = rk4_sens(X[k], U[k], p[k], P)
        b[k] = xn - X[k + 1]
    # [upstream D1] stage cost scaled by dt, terminal cost unscaled;
    # [upstream D7] Gauss-Newton Hessian V'WV with V = selection (blastermodel.py:247-254)
    Hu = np.tile(dt * P.R, (N, 1))
    Hx = np.vstack([np.tile(dt * P.Q, (N, 1)), P.Qt[None]])
    gu = dt * P.R * (U - yref[:N, nx:])
    gx = np.vstack([dt * P.Q * (X[:N] - yref[:N, :nx]), (P.Qt * (X[N] - yref[N, :nx]))[None]])
    # [upstream D2] lbx/ubx on stages 1..N-1 only; lbu/ubu on 0..N-1; none at N
    lbx = np.full((N + 1, nx), -np.inf)
    ubx = np.full((N + 1, nx), np.inf)
    lbx[1:N] = P.lbx - X[1:N]
    ubx[1:N] = P.ubx - X[1:N]
    # [upstream D3] stage-0 state pinned as equality (set(0,'lbx'/'ubx',x), simulation_blaster.py:60-61)
    return StageQP(A, B, b, Hu, Hx, gu, gx, P.lbu - U, P.ubu - U, lbx, ubx, x0 - X[0])


def qp_to_dense(qp: StageQP):
    """Stack into min 1/2 z'Hz + g'z s.t. Cz = c, lb<=z<=ub with
    z = [du_0, dx_1, du_1, dx_2, ..., du_{N-1}, dx_N]  (dx_0 eliminated)."""
    N, nx, nu = qp.A.shape[0], qp.A.shape[1], qp.B.shape[2]
    nz = nu + nx
    n = N * nz
    H = np.zeros(n)
    g = np.zeros(n)
    lb = np.zeros(n)
    ub = np.zeros(n)
    C = np.zeros((N * nx, n))
    c = np.zeros(N * nx)
    for k in range(N):
        o = k * nz
        H[o:o + nu] = qp.Hu[k]
        g[o:o + nu] = qp.gu[k]
        lb[o:o + nu] = qp.lbu[k]
        ub[o:o + nu] = qp.ubu[k]
        H[o + nu:o + nz] = qp.Hx[k + 1]
        g[o + nu:o + nz] = qp.gx[k + 1]
        lb[o + nu:o + nz] = qp.lbx[k + 1]
        ub[o + nu:o + nz] = qp.ubx[k + 1]
        r = slice(k * nx, (k + 1) * nx)
        # dx_{k+1} - B du_k - A dx_k = b_k
        C[r, o + nu:o + nz] = np.eye(nx)
        C[r, o:o + nu] = -qp.B[k]
        if k == 0:
            c[r] = qp.b[0] + qp.A[0] @ qp.dx0
        else:
            C[r, o - nx:o] = -qp.A[k]
            c[r] = qp.b[k]
    return H, g, C, c, lb, ub


def split_z(z, qp: StageQP):
    N, nx, nu = qp.A.shape[0], qp.A.shape[1], qp.B.shape[2]
    zz = z.reshape(N, nu + nx)
    du = zz[:, :nu].copy()
    dx = np.vstack([qp.dx0[None], zz[:, nu:]])
    return du, dx


# --------------------------------------------------------------------------
# A6: QP solve -- dense-KKT Mehrotra predictor-corrector interior point
# --------------------------------------------------------------------------
@dataclass
class IPMResult:
    z: np.ndarray
    pi: np.ndarray
    lam_l: np.ndarray
    lam_u: np.ndarray
    iters: int
    status: int  # 0 ok, 2 max-iter   [upstream D9]
    res: tuple


def ipm_dense(H, g, C, c, lb, ub, tol_stat=1e-6, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8, mu0=1e2, thr0=-0.5,
              max_iter=None, alpha_min=1e-8, z_fixed=None, strict=False) -> IPMResult:
    """Mehrotra predictor-corrector IPM, cold start [upstream D8], default tolerances =
    HPIPM's documented defaults (stationarity 1e-6, the rest 1e-8).  Same iteration
    (initial point, centering rule, single step length for primal and dual, step
    factor max(0.995, 1-mu_aff), stopping test) as the Riccati-based C oracle and the
    CUDA path, but every Newton system is solved on the *dense* KKT matrix with
    pivoted LU.

    The linear residuals (stationarity, dynamics, bound-slack) enter the stopping test
    through their exact-arithmetic values |r_0| * prod(1 - alpha_j): they are affine in
    the iterate and everything takes the same step (DESIGN.md "stopping test"): stationarity
    extrapolated from the start, equality and bound-slack residuals as measured on the previous
    iterate times (1 - alpha).  ``res`` also carries the explicitly evaluated norms for the tests.

    ``strict`` (mpcb_config.strict_reference): the reference stack's semantics -- the explicit norms
    (stationarity included) decide the stopping test, there is no early exit on diverging multipliers,
    the corrector solve gets one step of iterative refinement, and the iteration cap defaults to 500
    (blastermodel.py:279) instead of 60."""
    if max_iter is None:
        max_iter = 500 if strict else 60
    n, m = g.size, c.size
    il, iu = np.isfinite(lb), np.isfinite(ub)
    lbf = np.where(il, lb, 0.0)
    ubf = np.where(iu, ub, 0.0)
    z = np.zeros(n)
    pi = np.zeros(m)
    # slack floor: absolute (thr0 >= 0) or the fraction -thr0 of the box width (thr0 < 0)
    flo = thr0 if thr0 >= 0 else -thr0 * np.where(il & iu, ubf - lbf, 1.0)
    tl = np.where(il, np.maximum(z - lbf, flo), 1.0)
    tu = np.where(iu, np.maximum(ubf - z, flo), 1.0)
    ll = np.where(il, mu0 / tl, 0.0)
    lu = np.where(iu, mu0 / tu, 0.0)
    nb = int(il.sum() + iu.sum())
    K = np.zeros((n + m, n + m))
    K[:n, n:] = C.T
    K[n:, :n] = C
    status = 2
    res = None
    rg_est = None
    it = 0
    for it in range(max_iter):
        r_g = H * z + g + C.T @ pi - ll + lu
        r_b = C @ z - c
        r_dl = np.where(il, z - lbf - tl, 0.0)
        r_du = np.where(iu, ubf - z - tu, 0.0)
        mu = (ll @ tl + lu @ tu) / nb if nb else 0.0
        comp = max(np.max(ll * tl * il, initial=0.0), np.max(lu * tu * iu, initial=0.0))
        explicit = (np.abs(r_g).max(), np.abs(r_b).max(initial=0.0),
                    max(np.abs(r_dl).max(initial=0.0), np.abs(r_du).max(initial=0.0)))
        if rg_est is None:
            rg_est, rb_est, rd_est = explicit
        res = ((explicit if strict else (rg_est, rb_est, rd_est)) + (comp,)) + explicit
        if not np.isfinite(res[0] + res[1] + mu):
            status = 1
            break
        if not strict and mu > 1e2 * mu0:  # diverging multipliers = infeasible QP: stop early, same status as the min-step exit
            status = 3
            break
        if res[0] <= tol_stat and res[1] <= tol_eq and res[2] <= tol_ineq and comp <= tol_comp:
            status = 0
            break
        gam = ll / tl + lu / tu
        K[np.arange(n), np.arange(n)] = H + gam
        lu_piv = _lu_factor(K)

        def solve(r_ml, r_mu, refine=False):
            rhs = np.concatenate([-r_g - (r_ml + ll * r_dl) / tl + (r_mu + lu * r_du) / tu, -r_b])
            sol = _lu_solve(lu_piv, rhs)
            if refine:  # one step of iterative refinement (HPIPM: itref_corr_max), as the strict CUDA instantiation does
                sol = sol + _lu_solve(lu_piv, rhs - K @ sol)
            dz, dpi = sol[:n], sol[n:]
            dtl = np.where(il, dz + r_dl, 0.0)
            dtu = np.where(iu, -dz + r_du, 0.0)
            dll = np.where(il, -(r_ml + ll * dtl) / tl, 0.0)
            dlu = np.where(iu, -(r_mu + lu * dtu) / tu, 0.0)
            return dz, dpi, dtl, dtu, dll, dlu

        def max_step(dtl, dtu, dll, dlu):
            a = np.inf
            for v, dv, msk in ((tl, dtl, il), (tu, dtu, iu), (ll, dll, il), (lu, dlu, iu)):
                neg = msk & (dv < 0)
                if neg.any():
                    a = min(a, np.min(-v[neg] / dv[neg]))
            return a

        if nb:
            # predictor (affine scaling)
            dz, dpi, dtl, dtu, dll, dlu = solve(ll * tl, lu * tu)
            a_aff = min(1.0, max_step(dtl, dtu, dll, dlu))
            mu_aff = (((ll + a_aff * dll) * (tl + a_aff * dtl))[il].sum()
                      + ((lu + a_aff * dlu) * (tu + a_aff * dtu))[iu].sum()) / nb
            sigma = (mu_aff / mu) ** 3
            # corrector + centering
            dz, dpi, dtl, dtu, dll, dlu = solve(np.where(il, ll * tl + dll * dtl - sigma * mu, 0.0),
                                                np.where(iu, lu * tu + dlu * dtu - sigma * mu, 0.0), refine=strict)
            a = min(1.0, max(0.995, 1.0 - mu_aff) * max_step(dtl, dtu, dll, dlu))
        else:
            dz, dpi, dtl, dtu, dll, dlu = solve(np.zeros(n), np.zeros(n))
            a = 1.0
        z = z + a * dz
        pi = pi + a * dpi
        tl = tl + a * dtl
        tu = tu + a * dtu
        ll = ll + a * dll
        lu = lu + a * dlu
        rg_est *= (1.0 - a)
        rb_est = explicit[1] * (1.0 - a)  # measured on this iterate, then the exact-arithmetic decay of one step
        rd_est = explicit[2] * (1.0 - a)
        if not (a >= alpha_min):
            status = 3 if a == a else 1
            it += 1
            break
    if status == 2:
        it = max_iter
    return IPMResult(z, pi, np.where(il, ll, 0.0), np.where(iu, lu, 0.0), it, status, res)


def _lu_factor(K):
    import scipy.linalg as sla
    return sla.lu_factor(K, check_finite=False)


def _lu_solve(piv, rhs):
    import scipy.linalg as sla
    x = sla.lu_solve(piv, rhs, check_finite=False)
    return x


def kkt_certificate(H, g, C, c, lb, ub, z, pi, lam_l, lam_u):
    """Solver-independent optimality certificate: returns the inf-norms of
    (stationarity, equality residual, bound violation, complementarity,
    negative multipliers)."""
    il, iu = np.isfinite(lb), np.isfinite(ub)
    stat = np.abs(H * z + g + C.T @ pi - lam_l + lam_u).max()
    eq = np.abs(C @ z - c).max(initial=0.0)
    viol = max(np.max(np.where(il, lb - z, 0.0), initial=0.0), np.max(np.where(iu, z - ub, 0.0), initial=0.0), 0.0)
    comp = max(np.max(np.where(il, lam_l * np.abs(z - np.where(il, lb, 0)), 0.0), initial=0.0),
               np.max(np.where(iu, lam_u * np.abs(np.where(iu, ub, 0) - z), 0.0), initial=0.0))
    neg = max(np.max(-lam_l, initial=0.0), np.max(-lam_u, initial=0.0))
    return dict(stat=stat, eq=eq, viol=viol, comp=comp, neg=neg)


def explicit_kkt_residuals(P: "BlasterProblem", z, pi, tl, tu, ll, lu, lb, ub, g, BAt, b):
    """Explicit KKT residuals of a batch of stage-ordered QP iterates, e.g. what the CUDA path's mpcb_debug_qp exports:
    z, tl, tu, ll, lu, lb, ub, g [B,N+1,nz] with z_k = [du_k; dx_k]; pi [B,N+1,nx]; BAt [B,N,nz,nx] = [B_k'; A_k']; b [B,N,nx].
    Everything is evaluated from the data, nothing is taken from the solver's own bookkeeping.  Returns per-instance
    inf-norms: stat (Lagrangian gradient, Riccati sign convention H z + g - ll + lu + [B A]' pi_{k+1} - pi_k), eq
    (dynamics), ineq (bound-slack identities), viol (true bound violation of z), comp (max lam * t), neg (most negative
    slack or multiplier), and stat_comp [B,nz]: the stationarity norm per component of the stage variable."""
    z, pi, tl, tu, ll, lu, lb, ub, g, BAt, b = (np.asarray(a, dtype=np.float64) for a in (z, pi, tl, tu, ll, lu, lb, ub, g, BAt, b))
    N, nx, nu = P.N, P.nx, P.nu
    nz = nx + nu
    H = np.vstack([np.tile(np.concatenate([P.dt * P.R, P.dt * P.Q]), (N, 1)), np.concatenate([np.ones(nu), P.Qt])[None]])
    k = np.arange(N + 1)[:, None]
    j = np.arange(nz)[None, :]
    var = np.where(j < nu, k < N, k >= 1)
    hasb = np.where(j < nu, k < N, (k >= 1) & (k < N))
    r = H * z + g - ll + lu
    r[:, :N] += np.einsum("bkjc,bkc->bkj", BAt, pi[:, 1:])
    r[:, :, nu:] -= pi
    r = np.where(var, r, 0.0)
    eq = b + np.einsum("bkjc,bkj->bkc", BAt, z[:, :N]) - z[:, 1:, nu:]
    lbf, ubf = np.where(hasb, lb, 0.0), np.where(hasb, ub, 0.0)
    rd = np.where(hasb, np.maximum(np.abs(z - lbf - tl), np.abs(ubf - z - tu)), 0.0)
    viol = np.where(hasb, np.maximum(np.maximum(lbf - z, z - ubf), 0.0), 0.0)
    comp = np.where(hasb, np.maximum(ll * tl, lu * tu), 0.0)
    neg = np.where(hasb, np.maximum(np.maximum(-tl, -tu), np.maximum(-ll, -lu)), 0.0)
    B = z.shape[0]
    return dict(stat=np.abs(r).reshape(B, -1).max(1), eq=np.abs(eq).reshape(B, -1).max(1), ineq=rd.reshape(B, -1).max(1),
                viol=viol.reshape(B, -1).max(1), comp=comp.reshape(B, -1).max(1), neg=neg.reshape(B, -1).max(1),
                stat_comp=np.abs(r).max(1))


def multipliers_from_primal(H, g, C, lb, ub, z, act_tol=1e-7):
    """Recover (pi, lam_l, lam_u) for a primal candidate z by least squares on the
    stationarity condition with multipliers only on (near-)active bounds.  Used to
    certify primal solutions that come without duals (e.g. the GPU's dx/du)."""
    n = z.size
    act_l = np.isfinite(lb) & (z - lb <= act_tol)
    act_u = np.isfinite(ub) & (ub - z <= act_tol)
    r = H * z + g
    cols = [C.T]
    il = np.flatnonzero(act_l)
    iu = np.flatnonzero(act_u)
    El = np.zeros((n, il.size))
    El[il, np.arange(il.size)] = -1.0
    Eu = np.zeros((n, iu.size))
    Eu[iu, np.arange(iu.size)] = 1.0
    M = np.hstack(cols + [El, Eu])
    sol = np.linalg.lstsq(M, -r, rcond=None)[0]
    m = C.shape[0]
    pi = sol[:m]
    lam_l = np.zeros(n)
    lam_u = np.zeros(n)
    lam_l[il] = sol[m:m + il.size]
    lam_u[iu] = sol[m + il.size:]
    return pi, lam_l, lam_u


def solve_qp(qp: StageQP, **ipm_opts):
    H, g, C, c, lb, ub = qp_to_dense(qp)
    r = ipm_dense(H, g, C, c, lb, ub, **ipm_opts)
    du, dx = split_z(r.z, qp)
    return du, dx, r


# --------------------------------------------------------------------------
# A7/A8: SQP-RTI driver with persistent, un-shifted iterate
# --------------------------------------------------------------------------
def stage_cost(X, U, yref, P: BlasterProblem):
    """[upstream D10] get_cost(): sum dt*1/2|y-yref|_W^2 + 1/2|x_N-yref_N|_We^2."""
    yref = expand_yref(yref, P)
    nx = P.nx
    ex = X[:P.N] - yref[:P.N, :nx]
    eu = U - yref[:P.N, nx:]
    eN = X[P.N] - yref[P.N, :nx]
    return 0.5 * P.dt * ((P.Q * ex * ex).sum() + (P.R * eu * eu).sum()) + 0.5 * (P.Qt * eN * eN).sum()


class RTIOracle:
    """One instance of the reference's controller: ``solve`` = the six acados
    calls of simulation_blaster.py:60-89 (set lbx/ubx at 0, yref on all stages,
    p on stages 0..N-1, one SQP_RTI iteration, get(0,'u'))."""

    def __init__(self, P: BlasterProblem, **ipm_opts):
        self.P = P
        self.ipm_opts = ipm_opts
        self.strict = bool(ipm_opts.get("strict", False))
        self.reset()

    def reset(self, x_init=None, u_init=None):
        # [upstream D4] initial iterate all zero unless the caller sets one
        self.X = np.zeros((self.P.N + 1, self.P.nx))
        self.U = np.zeros((self.P.N, self.P.nu))
        if x_init is not None:
            self.X[:] = np.asarray(x_init, dtype=np.float64)
        if u_init is not None:
            self.U[:] = np.asarray(u_init, dtype=np.float64)
        self.last = None

    def solve(self, x0, yref, p=None):
        P = self.P
        qp = build_qp(self.X, self.U, np.asarray(x0, dtype=np.float64), yref, p, P)
        du, dx, r = solve_qp(qp, **self.ipm_opts)
        # FIXED_STEP, step length 1.0 (JSON globalization / nlp_solver_step_length):
        # full step, no SQP-level line search; [upstream D5] iterate kept, not shifted
        # a failed QP leaves the iterate untouched (same rule as the product); with the reference's semantics the last
        # interior-point iterate is applied when the iteration cap was hit, as acados does
        if r.status == 0 or (self.strict and r.status == 2):
            self.X = self.X + dx
            self.U = self.U + du
        self.last = (qp, r)
        self._yref = yref
        return self.U[0].copy(), self.X.copy(), self.U.copy(), r.status

    def cost(self):
        return stage_cost(self.X, self.U, self._yref, self.P)

    # ---- SQP to convergence (acados' SQP loop with nlp_solver_max_iter / nlp_solver_tol_*)
    def nlp_residuals(self, x0, yref, p=None, pi=None, lam_l=None, lam_u=None):
        """inf-norms (stat, eq, ineq, comp) at the stored iterate for the multipliers of the last QP (dense ordering of
        ``qp_to_dense``: pi over the N dynamics rows, lam over z = [du_0, dx_1, du_1, ..., dx_N])."""
        P = self.P
        qp = build_qp(self.X, self.U, np.asarray(x0, dtype=np.float64), yref, p, P)
        H, g, C, c, lb, ub = qp_to_dense(qp)
        n = g.size
        pi = np.zeros(c.size) if pi is None else pi
        lam_l = np.zeros(n) if lam_l is None else lam_l
        lam_u = np.zeros(n) if lam_u is None else lam_u
        il, iu = np.isfinite(lb), np.isfinite(ub)
        # at dz = 0 the QP gradient g is the NLP cost gradient and C'pi the dynamics term of the Lagrangian
        stat = np.abs(g + C.T @ pi - lam_l + lam_u).max()
        eq = max(np.abs(qp.b).max(), np.abs(qp.dx0).max())
        # lb / ub are bounds on the increment: lb = bound - value, so the violation of the iterate is max(lb, -ub, 0)
        ineq = max(np.max(np.where(il, lb, -np.inf), initial=0.0), np.max(np.where(iu, -ub, -np.inf), initial=0.0))
        comp = max(np.max(np.abs(lam_l * np.where(il, lb, 0.0)), initial=0.0), np.max(np.abs(lam_u * np.where(iu, ub, 0.0)), initial=0.0))
        return stat, eq, ineq, comp

    def sqp_solve(self, x0, yref, p=None, max_iter=100, tol=1e-6):
        """-> (u0, X, U, status, sqp_iters, qp_iters, res): status 0 converged, 2 max_iter, 4 QP failure, 1 NaN."""
        tol = np.broadcast_to(np.asarray(tol, dtype=np.float64), (4,))
        pi = lam_l = lam_u = None
        n_qp = qp_it = 0
        status = 2
        for it in range(max_iter + 1):
            res = self.nlp_residuals(x0, yref, p, pi, lam_l, lam_u)
            if not np.isfinite(res[0] + res[1]):
                status = 1
                break
            if all(r <= t for r, t in zip(res, tol)):
                status = 0
                break
            if it == max_iter:
                break
            _, _, _, st = self.solve(x0, yref, p)
            r = self.last[1]
            n_qp += 1
            qp_it += r.iters
            if not (st == 0 or (self.strict and st == 2)):
                status = 1 if st == 1 else 4
                break
            pi, lam_l, lam_u = r.pi, r.lam_l, r.lam_u
        return self.U[0].copy(), self.X.copy(), self.U.copy(), status, n_qp, qp_it, res


def closed_loop(P: BlasterProblem, x0, yref, p=None, steps=10, **ipm_opts):
    """simulation_blaster.py:56-105 without printing/plotting."""
    ctl = RTIOracle(P, **ipm_opts)
    pp = default_params() if p is None else np.asarray(p, dtype=np.float64)
    x = np.asarray(x0, dtype=np.float64).copy()
    simX = np.zeros((steps + 1, P.nx))
    simU = np.zeros((steps, P.nu))
    iters = []
    simX[0] = x
    for i in range(steps):
        u0, _, _, _ = ctl.solve(x, yref, p)
        iters.append(ctl.last[1].iters)
        x = plant_step(x, u0, pp if pp.ndim == 1 else pp[0], P)
        simU[i] = u0
        simX[i + 1] = x
    return simX, simU, iters

"""Seeded synthetic workloads for BASELINE.json's configs.

All inputs are generated on the host in FP64 with ``numpy.random.default_rng(seed)``
so the CPU baseline and the GPU path see bit-identical inputs.  Every x0 lies inside
the reference's state bounds (simulation_blaster.py:28-29).
"""
from __future__ import annotations

import numpy as np

NX_FULL, NU_FULL, NP = 17, 6, 25


def to_quat13(x0, yref):
    """QUAT13 version of a QUAD12 scenario: x = [p, q(w,x,y,z), v, omega] with q the quaternion of
    Rz(psi) Ry(theta) Rx(phi) (blastermodel.py:122), reference attitude q = [1,0,0,0], u_ref = 0."""
    x0 = np.asarray(x0, dtype=np.float64)
    hf, ht, hp = x0[:, 3] / 2, x0[:, 4] / 2, x0[:, 5] / 2
    cf, sf, ct, st, cp, sp = np.cos(hf), np.sin(hf), np.cos(ht), np.sin(ht), np.cos(hp), np.sin(hp)
    q = np.stack([cp * ct * cf + sp * st * sf, cp * ct * sf - sp * st * cf, cp * st * cf + sp * ct * sf,
                  sp * ct * cf - cp * st * sf], axis=1)
    x13 = np.concatenate([x0[:, 0:3], q, x0[:, 6:12]], axis=1)
    y = np.zeros(yref.shape[:-1] + (17,))
    y[..., 0:3] = yref[..., 0:3]
    y[..., 3] = 1.0
    return np.ascontiguousarray(x13), y


def default_params() -> np.ndarray:
    """blastermodel.py:280-282: POC Jacobians 0, T_blast = 2.2*9.81."""
    p = np.zeros(NP)
    p[24] = 2.2 * 9.81
    return p


def hover_to_setpoint():
    """Config 1: simulation_blaster.py:47-48."""
    x0 = np.zeros((1, NX_FULL))
    yref = np.zeros((1, NX_FULL + NU_FULL))
    yref[0, 2] = 3.5
    yref[0, 14] = 0.2
    return x0, yref


def random_setpoints(B: int, seed: int = 1234, nx: int = 17, nu: int = 6, alpha_max: float = 0.5):
    """Configs 2/4/5: randomised x0 inside the bounds and a random position set-point.
    ``alpha_max`` bounds the initial gimbal deflection; large deflections (0.5 rad) are fine for
    single solves but, with the +-5 deg/s swivel-rate and body-rate bounds, make the hard vx/vy <= 1 m/s
    state bounds unreachable a few control steps into a closed loop, so config 4 uses 0.05."""
    rng = np.random.default_rng(seed)
    x0 = np.zeros((B, NX_FULL))
    x0[:, 0:2] = rng.uniform(-1.0, 1.0, (B, 2))
    x0[:, 2] = rng.uniform(0.5, 4.0, B)
    x0[:, 3:5] = rng.uniform(-0.03, 0.03, (B, 2))
    x0[:, 5] = rng.uniform(-0.30, 0.30, B)
    x0[:, 6:9] = rng.uniform(-0.3, 0.3, (B, 3))
    x0[:, 9:12] = rng.uniform(-0.03, 0.03, (B, 3))
    x0[:, 12] = rng.uniform(0.0, alpha_max, B)
    x0[:, 13] = rng.uniform(-0.6 * alpha_max, 0.6 * alpha_max, B)
    yref = np.zeros((B, NX_FULL + NU_FULL))
    yref[:, 0:2] = rng.uniform(-1.2, 1.2, (B, 2))
    yref[:, 2] = rng.uniform(0.5, 4.5, B)
    if nx == NX_FULL:
        return x0, yref
    if nx == 13:
        return to_quat13(x0, yref)
    y = np.zeros((B, nx + nu))
    y[:, :nx] = yref[:, :nx]
    return np.ascontiguousarray(x0[:, :nx]), y


def lemniscate_tracking(B: int, N: int, dt: float = 1.0 / 30, seed: int = 2345, nx: int = 17, nu: int = 6,
                        amp_xy: float = 1.2, amp_z: float = 1.0, period: float = 2.0):
    """Config 3: per-stage yref[B, N+1, ny] on a figure-eight with random phase; x0 on the curve."""
    if nx == 13:
        return to_quat13(*lemniscate_tracking(B, N, dt, seed, 12, 4, amp_xy, amp_z, period))
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * np.pi, B)
    t = np.arange(N + 1) * dt
    w = 2 * np.pi / period
    yref = np.zeros((B, N + 1, nx + nu))
    yref[:, :, 0] = amp_xy * np.sin(w * t[None] + ph[:, None])
    yref[:, :, 1] = amp_xy * np.sin(2 * w * t[None] + ph[:, None])
    yref[:, :, 2] = 2.5 + amp_z * np.sin(w * t[None])
    x0 = np.zeros((B, nx))
    x0[:, 0:3] = yref[:, 0, 0:3]
    x0[:, 3:5] = rng.uniform(-0.05, 0.05, (B, 2))
    x0[:, 6:9] = rng.uniform(-0.3, 0.3, (B, 3))
    return x0, yref


def hover_trim(nu: int = 6, mass: float = 9.0, T_blast: float = 2.2 * 9.81) -> np.ndarray:
    """Per-rotor thrust that balances gravity with the nozzle pointing down (alpha = 0):
    sum(T) = M*9.81 - T_blast  (from blastermodel.py:163)."""
    u = np.zeros(nu)
    u[:4] = (mass * 9.81 - T_blast) / 4.0
    return u


def closed_loop_setpoints(B: int, seed: int = 3456, nx: int = 17, nu: int = 6):
    """Config 4 (closed-loop Monte-Carlo): the reference's hover-to-set-point manoeuvre
    (simulation_blaster.py:47-48) randomised -- vertical moves (|dz| <= 1.5 m) with small lateral
    offsets (<= 0.1 m) and small initial velocities.  With the reference's tight hard state
    bounds (|v| <= 1 m/s, body rates and swivel rates <= 5 deg/s) the attitude authority is so low
    that, at the 0.67 s horizon of N = 20, larger lateral moves make the closed loop oscillate with
    growing amplitude until the hard bounds become infeasible (the reference itself flies N = 60)."""
    rng = np.random.default_rng(seed)
    x0 = np.zeros((B, NX_FULL))
    x0[:, 0:2] = rng.uniform(-0.5, 0.5, (B, 2))  # well inside the +-1.5 m arena bounds
    x0[:, 2] = rng.uniform(1.0, 3.5, B)
    x0[:, 3:5] = rng.uniform(-0.02, 0.02, (B, 2))
    x0[:, 5] = rng.uniform(-0.30, 0.30, B)
    x0[:, 6:9] = rng.uniform(-0.05, 0.05, (B, 3))
    x0[:, 9:12] = rng.uniform(-0.02, 0.02, (B, 3))
    yref = np.zeros((B, NX_FULL + NU_FULL))
    yref[:, 0:2] = x0[:, 0:2] + rng.uniform(-0.1, 0.1, (B, 2))
    yref[:, 2] = np.clip(x0[:, 2] + rng.uniform(-1.5, 1.5, B), 0.5, 4.5)
    if nx == NX_FULL:
        return x0, yref
    if nx == 13:
        return to_quat13(x0, yref)
    y = np.zeros((B, nx + nu))
    y[:, :nx] = yref[:, :nx]
    return np.ascontiguousarray(x0[:, :nx]), y

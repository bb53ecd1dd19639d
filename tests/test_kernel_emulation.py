"""CPU tests of the CUDA kernel *source*: mpc_blaster_b200/csrc/*.cuh compiled with g++ under
-DMPCB_HOST_EMU (a warp = 32 fibers, tests/emu/warp_emu.h) and compared with the C oracle.
This exercises the warp-level logic (lane mapping, shuffles, shared-memory hand-offs) without a
GPU; the -m gpu tests check the same code on the device."""
import os

import numpy as np
import pytest

import emu_binding as eb
from mpc_blaster_b200 import scenarios as sc
from oracle import blaster_oracle as bo
from oracle import c_oracle as co

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("variant", [17, 12])
def test_emulated_kernels_match_oracle(variant):
    P = bo.canonical_problem(10, variant)
    x0, yref = sc.random_setpoints(2, seed=31, nx=P.nx, nu=P.nu)
    p = bo.default_params()
    for i in range(2):
        c = co.BatchRTI(P, 1, nthreads=1)
        X = np.zeros((P.N + 1, P.nx))
        U = np.zeros((P.N, P.nu))
        if i == 1:  # second instance starts from an initialised iterate
            X[:] = x0[i]
            U[:] = sc.hover_trim(P.nu)
            c.reset(x0[i:i + 1], sc.hover_trim(P.nu))
        x = x0[i].copy()
        for step in range(2):
            Xb, Ub = X.copy(), U.copy()
            st, it, BAt, b = eb.rti_solve(P, X, U, x, yref[i], p)
            for k in range(P.N):
                xn, A, B = co.rk4_sens(P, Xb[k], Ub[k], p)
                assert np.abs(BAt[k][P.nu:].T - A).max() < 1e-14 and np.abs(BAt[k][:P.nu].T - B).max() < 1e-14
                assert np.abs(b[k] - (xn - Xb[k + 1])).max() < 1e-14
            u0, Xc, Uc, stc = c.solve(x[None], yref[i], p)
            assert st == stc[0] == 0 and it == c.iters[0]
            assert np.abs(Xc[0] - X).max() < 1e-8 and np.abs(Uc[0] - U).max() < 1e-7
            xn = eb.plant_step(P, x, U[0], p)
            assert np.abs(xn - co.plant_step(P, x, U[0])[0]).max() < 1e-14
            x = xn


def test_emulated_kernel_per_stage_inputs():
    P = bo.canonical_problem(8)
    x0, yref = sc.lemniscate_tracking(1, 8)
    rng = np.random.default_rng(2)
    p = np.zeros((8, 25))
    p[:, :24] = 0.05 * rng.standard_normal(24)
    p[:, 24] = 21.0
    X = np.repeat(x0, 9, axis=0)
    U = np.tile(sc.hover_trim(), (8, 1))
    c = co.BatchRTI(P, 1, nthreads=1)
    c.reset(x0, sc.hover_trim())
    st, it, _, _ = eb.rti_solve(P, X, U, x0[0], yref[0], p)
    u0, Xc, Uc, stc = c.solve(x0, yref, p[None])
    assert st == stc[0] and it == c.iters[0]
    assert np.abs(Xc[0] - X).max() < 1e-8 and np.abs(Uc[0] - U).max() < 1e-7


def test_quaternion_helpers_match_reference_mathutils():
    """Device helpers restating utils/MathUtils.py:5-54 against values produced by the
    reference's own functions (tests/golden/make_golden.py)."""
    import ctypes as C
    g = np.load(os.path.join(G, "mathutils_golden.npz"))
    lib = eb.lib()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for i in range(g["q1"].shape[0]):
        q1, q2 = np.ascontiguousarray(g["q1"][i]), np.ascontiguousarray(g["q2"][i])
        out = np.zeros(4)
        lib.emu_quat_mul(dp(q1), dp(q2), dp(out))
        assert np.abs(out - g["prod"][i]).max() < 1e-15
        lib.emu_quat_inv(dp(q1), dp(out))
        assert np.array_equal(out, g["inv"][i])
        R = np.zeros(9)
        lib.emu_quat_to_rot(dp(q1), dp(R))
        assert np.abs(R.reshape(3, 3) - g["rot"][i]).max() < 1e-15
    # property from SURVEY section 4: quat2Rot(q(phi,theta,psi)) == Rz Ry Rx of the OCP
    lib.emu_euler_to_quat.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double)]
    q, R = np.zeros(4), np.zeros(9)
    lib.emu_euler_to_quat(0.13, -0.07, 0.31, dp(q))
    lib.emu_quat_to_rot(dp(q), dp(R))
    assert np.abs(R.reshape(3, 3) - bo._rot(0.13, -0.07, 0.31)).max() < 1e-15

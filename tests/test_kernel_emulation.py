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


@pytest.mark.parametrize("variant", [17, 12, 13])
def test_emulated_kernels_match_oracle(variant):
    P = bo.canonical_problem(10, variant)
    x0, yref = sc.random_setpoints(2, seed=31, nx=P.nx, nu=P.nu)
    p = bo.default_params()
    for i in range(2):
        c = co.BatchRTI(P, 1, nthreads=1)
        X = np.zeros((P.N + 1, P.nx))
        U = np.zeros((P.N, P.nu))
        if i == 1 or variant == 13:  # initialised iterate (always for QUAT13: the all-zero iterate is no quaternion)
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


def test_emulated_poc_generator_matches_oracle_and_reference_golden():
    """mpcb_poc.cuh compiled for the host against (a) the golden values produced by the reference's
    own Jacobian_POC_Solver.py / htm.py (tests/golden/make_poc_golden.py) and (b) the oracle.
    Forward differences with eps = 1e-6 amplify rounding differences by 1e6, hence 5e-6 on the
    Jacobians; the analytic mode is compared with the oracle's exact counterpart."""
    from oracle import poc_oracle as po
    g = np.load(os.path.join(G, "poc_golden.npz"))
    for i in range(g["euler"].shape[0]):
        e, m, p = g["euler"][i], g["motor"][i], g["position"][i]
        poc, Jm, Je, Jp, tf, st, p25 = eb.poc(e, m, p, mode=0)
        assert st == 0
        assert np.abs(poc - g["poc"][i]).max() < 1e-11 and abs(tf - g["t_flight"][i]) < 1e-13
        assert np.abs(Jm - g["J_mot"][i]).max() < 5e-6 and np.abs(Je - g["J_eul"][i]).max() < 5e-6
        assert np.abs(Jp - g["J_pos"][i]).max() < 5e-6
        # packing of simulation_blaster.py:67
        ref = np.concatenate([Jm.reshape(-1, order="F"), Je.reshape(-1, order="F"), Jp.reshape(-1, order="F"), [21.582]])
        assert np.array_equal(p25, ref)
        poc_a, Jm_a, Je_a, Jp_a, tf_a, st_a, _ = eb.poc(e, m, p, mode=1)
        o = po.analytic_jacobians(e, m, p)
        assert st_a == 0 and abs(poc_a[2]) < 1e-12
        assert np.abs(poc_a - o[0]).max() < 1e-12 and abs(tf_a - o[4]) < 1e-14
        assert np.abs(Jm_a - o[1]).max() < 1e-6 and np.abs(Je_a - o[2]).max() < 1e-6 and np.abs(Jp_a - o[3]).max() < 1e-6
        # the exact Jacobians differ from the reference's by what its |z| <= 1e-3 root tolerance leaves
        assert np.abs(Jm_a - Jm).max() < 2e-3 and np.abs(Je_a - Je).max() < 2e-3 and np.abs(Jp_a - Jp).max() < 2e-3


@pytest.mark.parametrize("variant,N,nb", [(17, 10, 4), (12, 10, 4), (17, 6, 3), (12, 5, 1), (12, 8, 11), (17, 5, 9), (13, 8, 6)])
def test_emulated_four_instances_per_warp_kernel_matches_oracle(variant, N, nb):
    """mpcb_qp8.cuh (four instances per warp, eight lanes each) compiled for the host: full and
    partly filled warps, two RTI steps (the four instances then differ in their IPM iteration
    counts, so groups finish at different times), against the C oracle.  nb > 4: the warp's groups
    are refilled from the work counter as their instances end (continuous batching)."""
    P = bo.canonical_problem(N, variant)
    x0, yref = sc.random_setpoints(nb, seed=31, nx=P.nx, nu=P.nu)
    p = bo.default_params()
    X = np.repeat(x0[:, None, :], N + 1, axis=1).copy()
    U = np.tile(sc.hover_trim(P.nu), (nb, N, 1)).copy()
    c = co.BatchRTI(P, nb, nthreads=1)
    c.reset(x0, sc.hover_trim(P.nu))
    x = x0.copy()
    for step in range(2):
        st, it = eb.rti_solve4(P, X, U, x, yref, p)
        u0, Xc, Uc, stc = c.solve(x, yref, p)
        assert (st == stc).all() and (st == 0).all() and (it == c.iters).all()
        assert np.abs(Xc - X).max() < 1e-8 and np.abs(Uc - U).max() < 1e-7
        x = co.plant_step(P, x, u0)


def test_emulated_kernel_reference_script_configuration():
    """The configuration simulation_blaster.py runs (N = 60, Tf = 2.0, parameters from the POC-Jacobian
    generator's reference mode, x0 = 0, set-point of :47-48, zero initial iterate) through the kernel
    source on the host: two control steps against the C oracle."""
    from oracle import poc_oracle as po
    P = bo.canonical_problem(60)
    _, J_mot, J_eul, J_pos = po.solve_jacobians([0, 0, 0], [0, 0], [0, 0, 4])
    p = bo.pack_params(J_mot, J_eul, J_pos, 2.2 * 9.81)
    x0, yref = bo.canonical_x0_yref()
    X, U = np.zeros((61, 17)), np.zeros((60, 6))
    c = co.BatchRTI(P, 1, nthreads=1)
    x = x0.copy()
    for step in range(2):
        st, it, _, _ = eb.rti_solve(P, X, U, x, yref, p)
        u0, Xc, Uc, stc = c.solve(x[None], yref, p)
        assert st == stc[0] == 0 and it == c.iters[0]
        assert np.abs(Xc[0] - X).max() < 1e-8 and np.abs(Uc[0] - U).max() < 1e-7
        x = co.plant_step(P, x, u0[0], p)[0]


_HYBRID_CASE = r"""
import sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import emu_binding as eb
from mpc_blaster_b200 import scenarios as sc
from oracle import blaster_oracle as bo
out = {}
# figure-eight tracking (state bounds active) and the hover-to-setpoint scenario, two warm-started steps each
for name, N, mk in (("track", 12, lambda P: sc.lemniscate_tracking(2, 12)), ("hover", 20, lambda P: tuple(a[None] for a in bo.canonical_x0_yref()))):
    P = bo.canonical_problem(N)
    x0, yref = mk(P)
    p = bo.default_params()
    for i in range(len(x0)):
        X = np.repeat(x0[i][None], N + 1, axis=0).copy()
        U = np.tile(sc.hover_trim(), (N, 1)).copy()
        x = x0[i].copy()
        for step in range(2):
            st, it, _, _ = eb.rti_solve(P, X, U, x, yref[i], p)
            out[f"{name}{i}_{step}"] = np.concatenate([[st, it], X.ravel(), U.ravel()])
            x = eb.plant_step(P, x, U[0], p)
np.savez(sys.argv[3], **out)
"""


def test_hybrid_factorisation_agrees_with_householder_only_build(tmp_path):
    """qp_kernel factorises in normal-equations form (the classical recursion on P_k, products on the emulated tensor-core
    fragments, input pivots only) while mu > MPCB_GRAM_MU = 3e-6 and by the Householder LQ afterwards (mpcb_qp.cuh).  The
    same kernel source built with -DMPCB_GRAM_MU=1e30 (LQ on every iteration) must take the same number of iterations and
    land on the same iterate; always-Gram (-DMPCB_GRAM_MU=-1) is what this guards against: it changes iteration counts and
    moves u by up to 8e-6.  Two more builds pin the pieces of the P form against each other: the
    17th column of P on the tensor-core tiles as well (-DMPCB_TAIL_COLUMN=0), and the round-1 factor form of the
    normal-equations iterations (-DMPCB_DMMA=0: CUDA-core Gram matrix, 23-pivot Cholesky, factor carried)."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    script = tmp_path / "case.py"
    script.write_text(_HYBRID_CASE)
    res = {}
    for tag, defs in (("hybrid", ""), ("lq", "MPCB_GRAM_MU=1e30"), ("notail", "MPCB_TAIL_COLUMN=0"), ("factor", "MPCB_DMMA=0")):
        out = tmp_path / f"{tag}.npz"
        env = dict(os.environ, MPCB_EMU_DEFINES=defs)
        subprocess.check_call([sys.executable, str(script), os.path.dirname(here), os.path.join(here, "emu"), str(out)], env=env)
        res[tag] = np.load(out)
    assert sorted(res["hybrid"].files) == sorted(res["lq"].files) and len(res["lq"].files) == 6
    for k, tag in ((k, tag) for k in res["lq"].files for tag in ("hybrid", "notail", "factor")):
        a, b = res[tag][k], res["lq"][k]
        assert a[0] == b[0] == 0 and a[1] == b[1], (k, tag)    # status, IPM iterations
        # measured: hybrid 2e-12 (CUDA-core products), 6e-11 (tensor-core products: another summation order); Gram on every iteration 1e-9 on these cases and 8e-6 on cold random set-points
        assert np.abs(a[2:] - b[2:]).max() < 5e-10, (k, np.abs(a[2:] - b[2:]).max())


def test_emulated_kernel_infeasible_instance_keeps_the_checkers_status():
    """An infeasible linearised QP (N = 40 random set-point, zero iterate): the multipliers diverge and the
    solve must end with the checker's status after the checker's number of iterations.  With the
    normal-equations factorisation allowed at mu > mu0 this instance broke down one iteration early with the
    QP-failure status instead of the min-step status -- hence the `mu <= mu0` guard in mpcb_qp.cuh."""
    N = 40
    P = bo.canonical_problem(N)
    x0, yref = sc.random_setpoints(40, seed=1234)
    p = bo.default_params()
    i = 19
    X, U = np.zeros((N + 1, P.nx)), np.zeros((N, P.nu))
    st, it, _, _ = eb.rti_solve(P, X, U, x0[i], yref[i], p)
    c = co.BatchRTI(P, 1, nthreads=1)
    _, _, _, stc = c.solve(x0[i:i + 1], yref[i], p)
    assert stc[0] != 0 and st == stc[0] and it == c.iters[0]


@pytest.mark.parametrize("variant", [17, 12])
def test_emulated_kernel_strict_reference_semantics_match_both_oracles(variant):
    """mpcb_config.strict_reference: the stopping test uses the residual norms the backward sweep evaluates explicitly on
    the iterate (stationarity included), as HPIPM's does, instead of their extrapolated values.  The strict instantiation
    of qp_solve_warp (host emulation), the C oracle (Riccati) and the NumPy oracle (dense KKT) must stop after the same
    number of iterations with the same status and the same iterate; the default rule set must reach the same primal
    solution (it may stop an iteration earlier or later)."""
    N = 10
    P = bo.canonical_problem(N, variant)
    x0, yref = sc.random_setpoints(3, seed=77, nx=P.nx, nu=P.nu)
    p = bo.default_params()
    for i in range(3):
        X = np.repeat(x0[i][None], N + 1, axis=0).copy()
        U = np.tile(sc.hover_trim(P.nu), (N, 1)).copy()
        Xd, Ud = X.copy(), U.copy()
        st, it, _, _ = eb.rti_solve(P, X, U, x0[i], yref[i], p, strict=True)
        c = co.BatchRTI(P, 1, nthreads=1, strict=True)
        c.reset(x0[i:i + 1], sc.hover_trim(P.nu))
        _, Xc, Uc, stc = c.solve(x0[i:i + 1], yref[i], p)
        o = bo.RTIOracle(P, strict=True)
        o.reset(x0[i], sc.hover_trim(P.nu))
        _, Xo, Uo, sto = o.solve(x0[i], yref[i])
        assert c.o.ipm_max_iter == 500 and c.o.strict == 1
        assert st == stc[0] == sto and it == c.iters[0] == o.last[1].iters, (st, stc, sto, it, c.iters, o.last[1].iters)
        if st in (0, 2):  # the step is applied on success and, under the reference's semantics, on max-iter
            assert np.abs(Xc[0] - X).max() < 1e-8 and np.abs(Uc[0] - U).max() < 1e-7
            assert np.abs(Xo - X).max() < 1e-8 and np.abs(Uo - U).max() < 1e-7
        # the explicit stationarity norm of the strict solve is what the test saw
        if st == 0:
            assert o.last[1].res[4] <= 1e-6 and o.last[1].res[5] <= 1e-8 and o.last[1].res[6] <= 1e-8
        std, itd, _, _ = eb.rti_solve(P, Xd, Ud, x0[i], yref[i], p)
        if st == 0 and std == 0:
            assert abs(it - itd) <= 2
            assert np.abs(Xd - X).max() < 1e-6 and np.abs(Ud[:, :4] - U[:, :4]).max() < 1e-5


def test_strict_reference_applies_the_last_iterate_on_max_iter():
    """acados takes HPIPM's last iterate when the QP solver returns max-iter; the default rule set leaves the iterate
    untouched on any failure.  A cap of 4 interior-point iterations forces status 2 in both modes."""
    N = 8
    P = bo.canonical_problem(N)
    x0, yref = sc.random_setpoints(1, seed=5)
    p = bo.default_params()
    X0 = np.repeat(x0[0][None], N + 1, axis=0)
    U0 = np.tile(sc.hover_trim(), (N, 1))
    X, U = X0.copy(), U0.copy()
    st, it, _, _ = eb.rti_solve(P, X, U, x0[0], yref[0], p, max_iter=4)
    assert st == 2 and it == 4 and np.array_equal(X, X0) and np.array_equal(U, U0)
    Xs, Us = X0.copy(), U0.copy()
    st, it, _, _ = eb.rti_solve(P, Xs, Us, x0[0], yref[0], p, max_iter=4, strict=True)
    c = co.BatchRTI(P, 1, nthreads=1, strict=True, max_iter=4)
    c.reset(x0, sc.hover_trim())
    _, Xc, Uc, stc = c.solve(x0, yref[0], p)
    assert st == stc[0] == 2 and it == c.iters[0] == 4
    assert np.abs(Us - U0).max() > 1e-3  # the step was taken
    assert np.abs(Xc[0] - Xs).max() < 1e-9 and np.abs(Uc[0] - Us).max() < 1e-8


def test_strict_reference_reaches_the_explicit_tolerances_with_state_bounds_active():
    """Config 1 (hover to set-point from the all-zero iterate): vz rides its bound on most stages from the first step.
    Without iterative refinement the explicitly evaluated stationarity norm of such solves stalls at 1e-5 .. 1e-3 (in the
    multipliers of the active bounds) and an explicit-norm test never succeeds; the strict instantiation refines the
    corrector solve once and must report success after the number of iterations the default rule set needs, with the
    same primal step."""
    N = 20
    P = bo.canonical_problem(N)
    x0, yref = bo.canonical_x0_yref()
    p = bo.default_params()
    Xs, Us = np.zeros((N + 1, P.nx)), np.zeros((N, P.nu))
    Xd, Ud = Xs.copy(), Us.copy()
    c = co.BatchRTI(P, 1, nthreads=1, strict=True)
    x = x0.copy()
    for step in range(4):
        c.X[0], c.U[0] = Xs, Us
        Xd[:], Ud[:] = Xs, Us
        st, it, _, _ = eb.rti_solve(P, Xs, Us, x, yref, p, strict=True)
        std, itd, _, _ = eb.rti_solve(P, Xd, Ud, x, yref, p)
        _, Xc, Uc, stc = c.solve(x[None], yref, p)
        assert st == stc[0] == 0 and it == c.iters[0], (step, st, stc, it, c.iters)
        assert std == 0 and abs(it - itd) <= 1, (step, it, itd)
        assert np.abs(Xc[0] - Xs).max() < 1e-8 and np.abs(Uc[0] - Us).max() < 1e-7
        assert np.abs(Xd - Xs).max() < 1e-7 and np.abs(Ud[:, :4] - Us[:, :4]).max() < 1e-6
        assert (np.abs(Xs[1:N, 8] - 1.0) < 1e-6).sum() >= 10   # vz <= 1 m/s is active on most stages
        x = eb.plant_step(P, x, Us[0], p)
    # bench-batch instances 1, 2, 4: the same explicit test WITHOUT the refinement step never succeeds (the interior
    # point runs on until the step length collapses), with it the emulated kernel and the oracle succeed together
    xb, yb = sc.random_setpoints(64, seed=1234)
    for i in (1, 2, 4):
        nr = co.BatchRTI(P, 1, nthreads=1, strict=True, itref=0, max_iter=60)
        nr.reset(xb[i:i + 1], sc.hover_trim())
        _, _, _, st0 = nr.solve(xb[i:i + 1], yb[i], p)
        assert st0[0] == 3
        wr = co.BatchRTI(P, 1, nthreads=1, strict=True)
        wr.reset(xb[i:i + 1], sc.hover_trim())
        _, Xc, Uc, st1 = wr.solve(xb[i:i + 1], yb[i], p)
        X = np.repeat(xb[i][None], N + 1, axis=0).copy()
        U = np.tile(sc.hover_trim(), (N, 1)).copy()
        st, it, _, _ = eb.rti_solve(P, X, U, xb[i], yb[i], p, strict=True)
        assert st == st1[0] == 0 and it == wr.iters[0]
        assert np.abs(Xc[0] - X).max() < 1e-8 and np.abs(Uc[0] - U).max() < 1e-7

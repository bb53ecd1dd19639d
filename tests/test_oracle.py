"""CPU tests of the oracle: pinned against fixtures generated from the reference's own
Python (tests/golden/make_golden.py) and cross-checked NumPy <-> C."""
import json
import os

import numpy as np
import pytest

from mpc_blaster_b200 import scenarios as sc
from oracle import blaster_oracle as bo
from oracle import c_oracle as co

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def dyn():
    return np.load(os.path.join(G, "dynamics_golden.npz"))


@pytest.fixture(scope="module")
def ocp():
    return json.load(open(os.path.join(G, "ocp_golden.json")))


def test_dynamics_match_reference_casadi_expression(dyn):
    """f, df/dx, df/du of the oracle == the reference's f_expl_expr (blastermodel.py:191-201)
    and its symbolic Jacobians, at 24 seeded points incl. the reference's start point."""
    P = bo.canonical_problem(20)
    for i in range(dyn["x"].shape[0]):
        x, u, p = dyn["x"][i], dyn["u"][i], dyn["p"][i]
        assert np.abs(bo.f17(x, u, p, P) - dyn["f"][i]).max() < 1e-12
        fx, fu = bo.jac17(x, u, p, P)
        assert np.abs(fx - dyn["fx"][i]).max() < 1e-12
        assert np.abs(fu - dyn["fu"][i]).max() < 1e-12
        assert np.abs(co.f(P, x, u, p) - dyn["f"][i]).max() < 1e-12


def test_jacobian_sparsity_is_the_reference_pattern(dyn):
    """59 structural nonzeros in df/dx, 32 in df/du."""
    assert int(dyn["fx_pattern"].sum()) == 59 and int(dyn["fu_pattern"].sum()) == 32
    P = bo.canonical_problem(20)
    fx, fu = bo.jac17(dyn["x"][3], dyn["u"][3], dyn["p"][3], P)
    assert ((fx != 0) <= (dyn["fx_pattern"] != 0)).all()
    assert ((fu != 0) <= (dyn["fu_pattern"] != 0)).all()


def test_problem_data_match_generateController_and_acados_dump(ocp):
    """Cost matrices, bounds, index sets and options of canonical_problem() equal both what the
    reference's generateController() sets (blastermodel.py:218-287) and the committed acados
    dump (src/scripts/acados_ocp_blasterModel.json)."""
    P = bo.canonical_problem(60)
    gc, dump = ocp["generateController"], ocp["acados_dump_extract"]
    W = np.array(gc["W"])
    assert np.count_nonzero(W - np.diag(np.diag(W))) == 0 and dump["W_offdiag_nnz"] == 0
    for diag in (np.diag(W), np.array(dump["W_diag"])):
        assert np.array_equal(diag, np.concatenate([P.Q, P.R]))
    for diag in (np.diag(np.array(gc["W_e"])), np.array(dump["W_e_diag"])):
        assert np.array_equal(diag, P.Qt)
    Vx, Vu = np.array(gc["Vx"]), np.array(gc["Vu"])
    assert np.array_equal(Vx, np.vstack([np.eye(17), np.zeros((6, 17))]))  # y = [x; u]
    assert np.array_equal(Vu, np.vstack([np.zeros((17, 6)), np.eye(6)]))
    assert np.array_equal(np.array(gc["Vx_e"]), np.eye(17))
    for src in (gc, dump):
        assert np.array_equal(src["lbx"], P.lbx) and np.array_equal(src["ubx"], P.ubx)
        assert np.array_equal(src["lbu"], P.lbu) and np.array_equal(src["ubu"], P.ubu)
        assert list(src["idxbx"]) == list(range(17)) and list(src["idxbu"]) == list(range(6))
        assert np.allclose(src["parameter_values"], bo.default_params(), rtol=1e-15)
    assert dump["idxbxe_0"] == list(range(17))  # stage-0 state is an equality
    d = dump["dims"]
    assert (d["nx"], d["nu"], d["np"], d["ny"], d["ny_e"], d["nbx"], d["nbu"], d["nbx_e"]) == (17, 6, 25, 23, 17, 17, 6, 0)
    so = dump["solver_options"]
    assert so["nlp_solver_type"] == "SQP_RTI" and so["integrator_type"] == "ERK" and so["hessian_approx"] == "GAUSS_NEWTON"
    assert so["qp_solver"] == "PARTIAL_CONDENSING_HPIPM" and so["qp_solver_cond_N"] == d["N"] == 60
    assert so["globalization"] == "FIXED_STEP" and so["nlp_solver_step_length"] == 1.0 and so["qp_solver_warm_start"] == 0
    assert so["levenberg_marquardt"] == 0.0
    assert len(dump["time_steps_unique"]) == 1 and abs(dump["time_steps_unique"][0] - P.dt) < 1e-15  # uniform dt = 1/30
    assert gc["solver_options"]["qp_solver_iter_max"] == 500


def test_param_packing_is_column_major():
    """simulation_blaster.py:67 (order='F') == blastermodel.py:203-210 (casadi reshape)."""
    Jm, Je, Jp = np.arange(6.).reshape(3, 2), 10 + np.arange(9.).reshape(3, 3), 20 + np.arange(9.).reshape(3, 3)
    p = bo.pack_params(Jm, Je, Jp, 7.0)
    assert p[1] == Jm[1, 0] and p[3] == Jm[0, 1] and p[6 + 1] == Je[1, 0] and p[15 + 3] == Jp[0, 1] and p[24] == 7.0
    P = bo.canonical_problem(5)
    x = np.zeros(17); x[6:9] = [1, 2, 3]
    u = np.zeros(6); u[4:6] = [0.5, -0.25]
    f = bo.f17(x, u, p, P)
    assert np.allclose(f[14:17], Jp @ x[6:9] + Jm @ u[4:6])


@pytest.mark.parametrize("variant", [17, 12])
def test_rk4_sensitivities_are_the_derivative_of_the_discrete_map(variant):
    P = bo.canonical_problem(10, variant)
    rng = np.random.default_rng(1)
    x0, _ = sc.random_setpoints(1, seed=5, nx=P.nx, nu=P.nu)
    x, u = x0[0], sc.hover_trim(P.nu) + rng.uniform(-1, 1, P.nu) * np.array([1, 1, 1, 1, .02, .02])[:P.nu]
    p = np.concatenate([0.2 * rng.standard_normal(24), [20.0]])
    xn, A, B = bo.rk4_sens(x, u, p, P)
    assert np.abs(xn - bo.plant_step(x, u, p, P)).max() < 1e-15
    e = 1e-6
    for j in range(P.nx):
        d = np.zeros(P.nx); d[j] = e
        assert np.abs((bo.plant_step(x + d, u, p, P) - bo.plant_step(x - d, u, p, P)) / (2 * e) - A[:, j]).max() < 1e-8
    for j in range(P.nu):
        d = np.zeros(P.nu); d[j] = e
        assert np.abs((bo.plant_step(x, u + d, p, P) - bo.plant_step(x, u - d, p, P)) / (2 * e) - B[:, j]).max() < 1e-8
    xc, Ac, Bc = co.rk4_sens(P, x, u, p)
    assert max(np.abs(xc - xn).max(), np.abs(Ac - A).max(), np.abs(Bc - B).max()) < 1e-14
    assert np.abs(co.plant_step(P, x, u, p)[0] - xn).max() < 1e-15


def test_dense_ipm_against_bounded_least_squares():
    """Box-constrained strictly convex QP without equalities: the IPM optimum equals
    scipy's BVLS solution of the equivalent least-squares problem."""
    from scipy.optimize import lsq_linear
    rng = np.random.default_rng(3)
    n = 12
    A = rng.standard_normal((30, n))
    bvec = rng.standard_normal(30)
    h = rng.uniform(0.5, 2.0, n)  # diagonal Hessian: scale columns so that A'A is not needed
    As = np.vstack([np.diag(np.sqrt(h))])
    t = rng.standard_normal(n) * 2
    lb, ub = -np.ones(n) * 0.7, np.ones(n) * 0.4
    # min 1/2 sum h (z - t)^2  ==  1/2 z'Hz + g'z with g = -h t
    r = bo.ipm_dense(h, -h * t, np.zeros((0, n)), np.zeros(0), lb, ub, tol_comp=1e-12)
    ref = lsq_linear(As, np.sqrt(h) * t, bounds=(lb, ub), method="bvls", tol=1e-14).x
    assert r.status == 0 and np.abs(r.z - ref).max() < 1e-9 and np.abs(r.z - np.clip(t, lb, ub)).max() < 1e-9


def test_qp_solution_carries_a_kkt_certificate():
    P = bo.canonical_problem(20)
    x0, yref = bo.canonical_x0_yref()
    qp = bo.build_qp(np.zeros((21, 17)), np.zeros((20, 6)), x0, yref, None, P)
    H, g, C, c, lb, ub = bo.qp_to_dense(qp)
    r = bo.ipm_dense(H, g, C, c, lb, ub, tol_comp=1e-12)
    cert = bo.kkt_certificate(H, g, C, c, lb, ub, r.z, r.pi, r.lam_l, r.lam_u)
    assert r.status == 0
    assert cert["stat"] < 1e-8 and cert["eq"] < 1e-10 and cert["viol"] < 1e-10 and cert["comp"] < 1e-9 and cert["neg"] == 0
    # the same primal, certified without the solver's own multipliers
    pi, ll, lu = bo.multipliers_from_primal(H, g, C, lb, ub, r.z)
    cert2 = bo.kkt_certificate(H, g, C, c, lb, ub, r.z, pi, ll, lu)
    assert cert2["stat"] < 1e-6 and cert2["neg"] < 1e-6


def test_hover_to_setpoint_behaviour():
    """Config 1 (simulation_blaster.py:47-48): step 0 saturates all four rotors at 65 N, then
    the climb rides the vz <= 1 m/s state bound with hover-trim thrust."""
    P = bo.canonical_problem(20)
    x0, yref = bo.canonical_x0_yref()
    simX, simU, iters = bo.closed_loop(P, x0, yref, steps=6)
    assert np.allclose(simU[0, :4], 65.0, atol=1e-5)
    assert abs(simU[3, :4].mean() - (9.0 * 9.81 - 2.2 * 9.81) / 4) < 1e-3  # hover trim 16.677 N per rotor
    assert abs(simX[3, 8] - 1.0) < 1e-6 and simX[:, 8].max() < 1.0 + 1e-7
    assert max(iters) < 25


@pytest.mark.parametrize("variant", [17, 12])
def test_c_oracle_follows_the_numpy_oracle(variant):
    """Same Mehrotra iteration, different linear algebra (Riccati-LQ vs dense LU): same
    iteration counts, same status, primal agreement far below the 1e-6 parity bound."""
    P = bo.canonical_problem(20, variant)
    B = 6
    x0, yref = sc.random_setpoints(B, seed=77, nx=P.nx, nu=P.nu)
    c = co.BatchRTI(P, B, nthreads=2)
    ctls = [bo.RTIOracle(P) for _ in range(B)]
    x = x0.copy()
    for step in range(3):
        u0, X, U, st = c.solve(x, yref)
        for i in range(B):
            ui, Xi, Ui, sti = ctls[i].solve(x[i], yref[i])
            assert sti == st[i] == 0 and ctls[i].last[1].iters == c.iters[i]
            assert np.abs(Ui - U[i]).max() < 1e-7 and np.abs(Xi - X[i]).max() < 1e-8
        x = co.plant_step(P, x, u0)


def test_tight_tolerance_converges_to_the_unique_optimum():
    """Driving the complementarity tolerance from HPIPM's 1e-8 to 1e-12 moves the well-determined
    part of the solution (states, thrusts of the first stage) by < 1e-4: u0 is insensitive."""
    P = bo.canonical_problem(20)
    x0, yref = sc.random_setpoints(4, seed=5)
    a = co.BatchRTI(P, 4, tol_comp=1e-8)
    b = co.BatchRTI(P, 4, tol_comp=1e-12)
    for o in (a, b):
        o.reset(x0, sc.hover_trim())
    ua, Xa, _, sa = a.solve(x0, yref)
    ub, Xb, _, sb = b.solve(x0, yref)
    assert (sa == 0).all() and (sb == 0).all() and (b.iters >= a.iters).all()
    assert np.abs(ua[:, :4] - ub[:, :4]).max() < 1e-4 and np.abs(Xa - Xb).max() < 1e-4


def test_infeasible_qp_is_reported_not_hidden():
    P = bo.canonical_problem(20)
    x0, yref = sc.random_setpoints(2, seed=15)
    x0[1, 6] = 3.0  # vx = 3 m/s against a +-1 m/s bound on stage 1
    c = co.BatchRTI(P, 2)
    c.reset(x0, sc.hover_trim())
    _, _, _, st = c.solve(x0, yref)
    assert st[0] == 0 and st[1] in (2, 3, 4)
    ctl = bo.RTIOracle(P)
    ctl.reset(x0[1], sc.hover_trim())
    assert ctl.solve(x0[1], yref[1])[3] == st[1]


def test_cost_matches_definition():
    P = bo.canonical_problem(5)
    rng = np.random.default_rng(0)
    X, U, yref = rng.standard_normal((6, 17)), rng.standard_normal((5, 6)), rng.standard_normal(23)
    ref = sum(0.5 * P.dt * (P.Q @ (X[k] - yref[:17]) ** 2 + P.R @ (U[k] - yref[17:]) ** 2) for k in range(5))
    ref += 0.5 * P.Qt @ (X[5] - yref[:17]) ** 2
    assert abs(bo.stage_cost(X, U, yref, P) - ref) < 1e-9 * abs(ref)


def test_hover_closed_loop_golden_regression():
    """Config 1: 200 closed-loop steps of the reference's scenario against the committed golden
    trajectory (tests/golden/make_closed_loop_golden.py; our oracle's own output, i.e. a regression
    pin).  The rigid-body states and thrusts must reproduce tightly; the swivel rates / gimbal
    angles are only weakly determined and get a loose bound."""
    g = np.load(os.path.join(G, "hover_closed_loop_golden.npz"))
    P = bo.canonical_problem(20)
    c = co.BatchRTI(P, 1, nthreads=1)
    x = g["simX"][0][None].copy()
    for s in range(60):
        u0, X, U, st = c.solve(x, g["yref"])
        assert st[0] == 0
        assert np.abs(u0[0, :4] - g["simU"][s, :4]).max() < 1e-6
        x = co.plant_step(P, x, u0)
        assert np.abs(x[0, :12] - g["simX"][s + 1, :12]).max() < 1e-7
        assert np.abs(x[0] - g["simX"][s + 1]).max() < 1e-3
    assert abs(g["simX"][-1][2] - 3.5) < 0.02 and g["simX"][:, 8].max() < 1.0 + 1e-6  # reaches z = 3.5 riding vz <= 1


def test_poc_oracle_reproduces_the_reference_jacobian_poc_solver():
    """oracle/poc_oracle.py against tests/golden/poc_golden.npz, i.e. against outputs of the
    reference's own Jacobian_POC_Solver.py and htm.py (run by tests/golden/make_poc_golden.py):
    homogeneous transforms, jet initial state, time of flight, POC and the three Jacobians for 12
    poses, and the Jacobians of the reference's initialise() call (simulation_blaster.py:37-39)."""
    from oracle import poc_oracle as po
    g = np.load(os.path.join(G, "poc_golden.npz"))
    V, c = float(g["stream_velocity"]), float(g["M_c"])
    for i in range(g["euler"].shape[0]):
        e, m, p = g["euler"][i], g["motor"][i], g["position"][i]
        assert np.abs(po.T_b_s2(*m) - g["T_b_s2"][i]).max() < 1e-15
        assert np.abs(po.T_w_b(*e, p) - g["T_w_b"][i]).max() < 1e-15
        x0 = po.init_conditions(e, m, p, V)
        assert np.abs(x0 - g["x_init"][i]).max() < 1e-12
        assert abs(po.time_of_flight(x0, c) - g["t_flight"][i]) < 1e-14
        poc, Jm, Je, Jp = po.solve_jacobians(e, m, p, V, c)
        assert np.abs(poc - g["poc"][i]).max() < 1e-11
        assert np.abs(Jm - g["J_mot"][i]).max() < 5e-6 and np.abs(Je - g["J_eul"][i]).max() < 5e-6
        assert np.abs(Jp - g["J_pos"][i]).max() < 5e-6
    _, Jm, Je, Jp = po.solve_jacobians([0, 0, 0], [0, 0], [0, 0, 4], V, c)
    assert np.abs(Jm - g["init_J_mot"]).max() < 5e-6 and np.abs(Je - g["init_J_eul"]).max() < 5e-6 and np.abs(Jp - g["init_J_pos"]).max() < 5e-6
    # the ndarray call pattern of the reference accumulates the position perturbations (documented, not a target)
    acc = np.cumsum(g["J_pos"][0], axis=1)
    assert np.abs(acc - g["J_pos_ndarray_call"]).max() < 1e-3


def test_quat13_model_is_the_euler_model_in_quaternion_coordinates():
    """QUAT13 exists nowhere in the reference, so it is pinned to what does: at
    q = q(phi, theta, psi) its rotation matrix is the OCP's Rz Ry Rx (blastermodel.py:122) through
    MathUtils.quat2Rot, its translational and angular accelerations equal QUAD12's, its qdot is the
    Euler-rate field pushed through dq/d(euler); Jacobians against central differences; the C oracle
    against the NumPy one; |q| is an invariant of the continuous dynamics."""
    P13, P12 = bo.canonical_problem(6, 13), bo.canonical_problem(6, 12)
    rng = np.random.default_rng(0)
    p = bo.default_params()
    eps = 1e-6
    for _ in range(6):
        x12 = np.concatenate([rng.uniform(-1, 1, 3), rng.uniform(-0.3, 0.3, 3), rng.uniform(-1, 1, 3), rng.uniform(-0.2, 0.2, 3)])
        u = rng.uniform(5, 40, 4)
        x13 = bo.x12_to_x13(x12)
        assert abs(np.linalg.norm(x13[3:7]) - 1) < 1e-14
        assert np.abs(bo.quat_to_rot(x13[3:7]) - bo._rot(*x12[3:6])).max() < 1e-14
        f12, f13 = bo.f(x12, u, p, P12), bo.f(x13, u, p, P13)
        assert np.abs(f12[0:3] - f13[0:3]).max() < 1e-14 and np.abs(f12[6:12] - f13[7:13]).max() < 1e-12
        Jq = np.stack([(bo.euler_to_quat(*(x12[3:6] + eps * np.eye(3)[i])) - bo.euler_to_quat(*(x12[3:6] - eps * np.eye(3)[i]))) / (2 * eps)
                       for i in range(3)], axis=1)
        assert np.abs(Jq @ f12[3:6] - f13[3:7]).max() < 1e-8
        assert abs(x13[3:7] @ f13[3:7]) < 1e-15                               # d|q|^2/dt = 0
        fx, fu = bo.jac(x13, u, p, P13)
        fxn = np.stack([(bo.f(x13 + eps * np.eye(13)[i], u, p, P13) - bo.f(x13 - eps * np.eye(13)[i], u, p, P13)) / (2 * eps) for i in range(13)], axis=1)
        fun = np.stack([(bo.f(x13, u + eps * np.eye(4)[i], p, P13) - bo.f(x13, u - eps * np.eye(4)[i], p, P13)) / (2 * eps) for i in range(4)], axis=1)
        assert np.abs(fx - fxn).max() < 1e-7 and np.abs(fu - fun).max() < 1e-8
        assert np.abs(co.f(P13, x13, u, p) - f13).max() < 1e-13
        xn, A, B = co.rk4_sens(P13, x13, u, p)
        xn2, A2, B2 = bo.rk4_sens(x13, u, p, P13)
        assert np.abs(xn - xn2).max() < 1e-13 and np.abs(A - A2).max() < 1e-13 and np.abs(B - B2).max() < 1e-13
    # one RTI iteration: dense-KKT NumPy oracle against the Riccati C oracle
    x0, yref = sc.random_setpoints(4, seed=17, nx=13, nu=4)
    trim = sc.hover_trim(4)
    c = co.BatchRTI(P13, 4, nthreads=1)
    c.reset(x0, trim)
    u0, X, U, st = c.solve(x0, yref)
    for i in range(4):
        o = bo.RTIOracle(P13)
        o.reset(x0[i], trim)
        uo, Xo, Uo, so = o.solve(x0[i], yref[i])
        assert so == st[i] == 0
        assert np.abs(Uo - U[i]).max() < 1e-7 and np.abs(Xo - X[i]).max() < 1e-8


def test_sqp_to_convergence_c_oracle_matches_numpy_oracle():
    """SQP to convergence with the options the reference's dump carries (nlp_solver_tol_* = 1e-6,
    nlp_solver_max_iter = 100, acados_ocp_blasterModel.json solver_options).  The C restatement (Riccati IPM) and the
    NumPy one (dense KKT) run the same loop -- linearise, NLP residuals with the last QP's multipliers, stop or solve the
    QP and take the full step -- and must agree on the number of QPs, the status and the converged iterate; the
    converged iterate must satisfy the NLP KKT conditions (the residuals are the certificate)."""
    from oracle import c_oracle as co
    from mpc_blaster_b200 import scenarios as sc
    N = 10
    P = bo.canonical_problem(N)
    x0, yref = sc.random_setpoints(3, seed=11)
    trim = sc.hover_trim()
    c = co.BatchRTI(P, 3, nthreads=1, strict=True, max_iter=60)  # the product solves the QPs of this mode with the refined rule set
    c.reset(x0, trim)
    u0, Xc, Uc, st, n_qp, qp_it, res = c.sqp_solve(x0, yref, max_iter=100, tol=1e-6)
    assert (st == 0).all() and (n_qp >= 2).all() and (n_qp < 30).all() and (res <= 1e-6).all(), (st, n_qp, res)
    for i in range(3):
        o = bo.RTIOracle(P, strict=True, max_iter=60)
        o.reset(x0[i], trim)
        _, Xo, Uo, sto, n_o, it_o, res_o = o.sqp_solve(x0[i], yref[i], max_iter=100, tol=1e-6)
        assert sto == 0 and n_o == n_qp[i] and it_o == qp_it[i], (i, sto, n_o, n_qp[i], it_o, qp_it[i])
        assert np.abs(Xo - Xc[i]).max() < 1e-6 and np.abs(Uo - Uc[i]).max() < 1e-6
        assert max(res_o) <= 1e-6 and np.abs(np.array(res_o) - res[i]).max() < 1e-7
    # an iteration cap below what convergence needs: status 2, exactly that many QPs, residuals above the tolerance
    c.reset(x0, trim)
    _, _, _, st2, n2, _, res2 = c.sqp_solve(x0, yref, max_iter=2, tol=1e-6)
    assert (st2 == 2).all() and (n2 == 2).all() and (res2.max(axis=1) > 1e-6).all()
    # one SQP iteration from the same start is the RTI step
    c.reset(x0, trim)
    _, X1, U1, _, _, _, _ = c.sqp_solve(x0, yref, max_iter=1, tol=0.0)
    r = co.BatchRTI(P, 3, nthreads=1, strict=True, max_iter=60)
    r.reset(x0, trim)
    _, Xr, Ur, _ = r.solve(x0, yref)
    assert np.array_equal(X1, Xr) and np.array_equal(U1, Ur)


def test_explicit_kkt_residual_evaluation_agrees_with_the_dense_certificate():
    """oracle.explicit_kkt_residuals evaluates the KKT residuals of a stage-ordered iterate (the layout mpcb_debug_qp
    exports from the GPU); on the NumPy oracle's own solution, re-packed into that layout, it must reproduce the dense
    certificate (kkt_certificate) of the same point -- so the GPU test's checker is itself checked."""
    from mpc_blaster_b200 import scenarios as sc
    N = 6
    P = bo.canonical_problem(N)
    x0, yref = sc.random_setpoints(1, seed=3)
    o = bo.RTIOracle(P)
    o.reset(x0[0], sc.hover_trim())
    o.solve(x0[0], yref[0])
    qp, r = o.last
    H, g, C, c, lb, ub = bo.qp_to_dense(qp)
    cert = bo.kkt_certificate(H, g, C, c, lb, ub, r.z, r.pi, r.lam_l, r.lam_u)
    nx, nu = P.nx, P.nu
    nz = nx + nu
    z = np.zeros((1, N + 1, nz)); pi = np.zeros((1, N + 1, nx)); ll = np.zeros_like(z); lu = np.zeros_like(z)
    lbs = np.full_like(z, -np.inf); ubs = np.full_like(z, np.inf); gs = np.zeros_like(z)
    zz, l_l, l_u = r.z.reshape(N, nz), r.lam_l.reshape(N, nz), r.lam_u.reshape(N, nz)
    lbd, ubd, gd = lb.reshape(N, nz), ub.reshape(N, nz), g.reshape(N, nz)
    z[0, 0, nu:] = qp.dx0
    BAt = np.zeros((1, N, nz, nx)); bb = np.zeros((1, N, nx))
    for k in range(N):
        for dst, src in ((z, zz), (ll, l_l), (lu, l_u), (lbs, lbd), (ubs, ubd), (gs, gd)):
            dst[0, k, :nu] = src[k, :nu]
            dst[0, k + 1, nu:] = src[k, nu:]
        pi[0, k + 1] = -r.pi[k * nx:(k + 1) * nx]  # the dense ordering carries the opposite sign of the Riccati one
        BAt[0, k, :nu], BAt[0, k, nu:], bb[0, k] = qp.B[k].T, qp.A[k].T, qp.b[k]
    fin_l, fin_u = np.isfinite(lbs), np.isfinite(ubs)
    tl = np.where(fin_l, z - np.where(fin_l, lbs, 0.0), 1.0)
    tu = np.where(fin_u, np.where(fin_u, ubs, 0.0) - z, 1.0)
    out = bo.explicit_kkt_residuals(P, z, pi, tl, tu, ll, lu, lbs, ubs, gs, BAt, bb)
    assert abs(out["stat"][0] - cert["stat"]) < 1e-15 and abs(out["comp"][0] - cert["comp"]) < 1e-18
    assert out["eq"][0] < 1e-14 and out["viol"][0] == 0.0 and out["ineq"][0] == 0.0 and out["neg"][0] <= 0.0
    assert out["stat_comp"].shape == (1, nz) and out["stat_comp"].max() == out["stat"][0]
    # a perturbed multiplier shows up in the stationarity norm of exactly its component
    ll2 = ll.copy()
    ll2[0, 2, 1] += 1e-3
    out2 = bo.explicit_kkt_residuals(P, z, pi, tl, tu, ll2, lu, lbs, ubs, gs, BAt, bb)
    assert abs(out2["stat_comp"][0, 1] - 1e-3) < 1e-9 and out2["stat_comp"][0, 0] == out["stat_comp"][0, 0]

"""GPU tests added in round 2: measured (not predicted) KKT residuals of the CUDA path's solutions, full-size parity
for BASELINE.json configs 3, 4 and 5, SQP to convergence, the reference-semantics switch, and the sharded ==
unsharded identity on real GPUs.  Everything goes through the C ABI (libmpcb.so via ctypes); the oracle is the checker.

Tests that measure write a small JSON report under gpurun_out/ (merged back by gpurun); the numbers quoted in
BASELINE.md come from those reports.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from mpc_blaster_b200 import scenarios as sc
from oracle import blaster_oracle as bo
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu
TOL = float(os.environ.get("MPCB_TEST_TOL", "1e-6"))  # FP64 parity bound of north_star (the override is for margin measurements only: tools/ab.py runs)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mpc(N, B, variant=17, **kw):
    from mpc_blaster_b200 import BlasterMPC
    return BlasterMPC.canonical(N=N, batch=B, variant=variant, **kw)


def _report(name, obj):
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, name), "w") as f:
            json.dump(obj, f, indent=1)
    except OSError:
        pass


def _kkt(mpc, P, B):
    d = {k: v.cpu().numpy() for k, v in mpc.debug_qp(B).items()}
    return bo.explicit_kkt_residuals(P, d["z"], d["pi"], d["tl"], d["tu"], d["ll"], d["lu"], d["lb"], d["ub"], d["g"], d["BAt"], d["b"])


def _quant(a):
    a = np.asarray(a, dtype=np.float64)
    return {"max": float(a.max()), "p99": float(np.percentile(a, 99)), "p50": float(np.percentile(a, 50))}


# --------------------------------------------------------------------------- measured KKT residuals
def _kkt_scenario(scenario):
    if scenario.startswith("bench"):
        N, B, kw = 20, 1024, {}
        x0, yref = sc.random_setpoints(B, seed=1234)
        p = None
    elif scenario == "tracking40_qp8":
        N, B, kw = 40, 512, dict(qp8_batch=1)
        x0, yref = sc.lemniscate_tracking(B, N)
        p = None
    else:
        from mpc_blaster_b200 import JacobianPOCSolver
        N, B, kw = 60, 64, {}
        x0, yref = sc.closed_loop_setpoints(B, seed=60)
        gen = JacobianPOCSolver(150, 1, 0.000015)
        gen.initialise()
        p = bo.pack_params(*gen.getJacobians(), 2.2 * 9.81)
    return N, B, kw, bo.canonical_problem(N), x0, yref, p


def _kkt_report(scenario, strict, mpc, P, B, ok, r):
    names = [f"u{j}" for j in range(P.nu)] + [f"x{i}" for i in range(P.nx)]
    rep = {"scenario": scenario, "strict_reference": bool(strict), "B": B, "N": P.N, "converged_frac": float(ok.mean()),
           "mean_ipm_iters": float(mpc.iters.double().mean()), "max_ipm_iters": int(mpc.iters.max()),
           "tolerances": {"stat": 1e-6, "eq": 1e-8, "ineq": 1e-8, "comp": 1e-8}}
    for k in ("stat", "eq", "ineq", "viol", "comp", "neg"):
        rep[k] = _quant(r[k][ok])
    rep["stat_frac_within_1e-6"] = float((r["stat"][ok] <= 1e-6).mean())
    rep["stat_max_by_component"] = {n: float(v) for n, v in zip(names, r["stat_comp"][ok].max(0))}
    _report(f"r02_kkt_{scenario}{'_strict' if strict else ''}.json", rep)
    return rep


@pytest.mark.parametrize("scenario", ["bench", "bench_zero_iterate", "tracking40_qp8", "script60"])
def test_explicit_kkt_residuals_of_gpu_solutions(cuda_device, scenario):
    """The stopping test of the kernels tracks the three linear residuals through their exact-arithmetic decay
     and confirms two of them explicitly.  Here all four KKT residuals of the interior-point
    iterate the GPU ended with are MEASURED: evaluated from the exported data (mpcb_debug_qp) by the NumPy oracle --
    nothing is taken from the solver's bookkeeping -- on the 1,024 bench instances (initialised and zero iterate), an
    N = 40 tracking batch with active state bounds on the four-instances-per-warp kernel, and the reference script's
    N = 60 configuration with its POC Jacobians.

    Asserted at HPIPM's default tolerances for every instance that reports success: dynamics 1e-8, bound-slack
    identities 1e-8, no bound violation, positivity (to the rounding of a step that lands on the boundary),
    complementarity 1e-8.  The explicit STATIONARITY norm is where the default rule set and the reference part: without
    iterative refinement the Riccati solve leaves up to ~1e-3 of it in the multipliers of active bounds (the primal
    step is accurate: the strict solve below, whose explicit norms all pass, lands on the same primal point to 1e-6), so
    here it is only bounded loosely and reported per component; BASELINE.md quotes the numbers, and
    test_strict_reference_meets_every_explicit_kkt_tolerance asserts HPIPM's 1e-6 for the refined solve."""
    N, B, kw, P, x0, yref, p = _kkt_scenario(scenario)
    mpc = _mpc(N, B, **kw)
    if scenario != "bench_zero_iterate":
        mpc.reset(x0, sc.hover_trim())
    u0, X, U, st = mpc.solve(x0, yref, p)
    ok = st.cpu().numpy() == 0
    assert ok.mean() > (0.9 if scenario != "bench_zero_iterate" else 0.5), ok.mean()
    r = _kkt(mpc, P, B)
    rep = _kkt_report(scenario, False, mpc, P, B, ok, r)
    assert r["eq"][ok].max() <= 1e-8, rep["eq"]
    assert r["ineq"][ok].max() <= 1e-8 and r["viol"][ok].max() <= 1e-8, (rep["ineq"], rep["viol"])
    assert r["comp"][ok].max() <= 1e-8 and r["neg"][ok].max() <= 1e-20, (rep["comp"], rep["neg"])
    assert np.isfinite(r["stat"][ok]).all()
    # measured in round 2: max 9.5e-4 / 82 % within 1e-6 on the bench batch, 3.8e-2 / 10 % on the N = 40
    # tracking batch (state bounds active on most stages), 4.2e-7 / 100 % on the N = 60 script configuration
    assert r["stat"][ok].max() <= 0.2, rep["stat"]


@pytest.mark.parametrize("scenario", ["bench", "tracking40_qp8", "script60"])
def test_strict_reference_meets_every_explicit_kkt_tolerance(cuda_device, scenario):
    """"Matched KKT tolerance", measured: with mpcb_config.strict_reference the kernel tests the explicitly evaluated
    norms and refines the corrector solve once (HPIPM's itref_corr); the exported iterate of every instance that reports
    success must then satisfy ALL four of HPIPM's default tolerances when re-evaluated by the NumPy oracle, in (almost)
    the number of iterations the default rule set predicted, and the default solve's primal solution must be the same
    point at north_star's 1e-6."""
    N, B, kw, P, x0, yref, p = _kkt_scenario(scenario)
    kw = dict(kw)
    kw.pop("qp8_batch", None)  # strict solves always run the one-instance kernel
    strict = _mpc(N, B, strict_reference=True, **kw)
    dflt = _mpc(N, B, **kw)
    for m in (strict, dflt):
        m.reset(x0, sc.hover_trim())
    us, Xs, Us, sts = strict.solve(x0, yref, p)
    ud, Xd, Ud, std = dflt.solve(x0, yref, p)
    ok = sts.cpu().numpy() == 0
    both = ok & (std.cpu().numpy() == 0)
    assert ok.mean() > 0.9 and both.mean() > 0.9
    r = _kkt(strict, P, B)
    rep = _kkt_report(scenario, True, strict, P, B, ok, r)
    assert r["stat"][ok].max() <= 1e-6, rep["stat"]
    assert r["eq"][ok].max() <= 1e-8 and r["ineq"][ok].max() <= 1e-8 and r["viol"][ok].max() <= 1e-8
    assert r["comp"][ok].max() <= 1e-8 and r["neg"][ok].max() <= 1e-20
    di = (strict.iters - dflt.iters).cpu().numpy()[both]
    assert np.abs(di).max() <= 2 and (di == 0).mean() > 0.9, (di.min(), di.max(), (di == 0).mean())
    tb = torch.as_tensor(both, device="cuda")
    dT = float((Us[tb][..., :4] - Ud[tb][..., :4]).abs().max())   # thrusts
    dS = float((Us[tb][..., 4:] - Ud[tb][..., 4:]).abs().max())   # swivel rates: curvature dt * 1e-5, determined to tol / 3.3e-7 only
    dX = float((Xs[tb] - Xd[tb]).abs().max())
    _report(f"r02_strict_vs_default_{scenario}.json", {"scenario": scenario, "B": B, "N": N, "max_d_thrust": dT, "max_d_swivel_rate": dS, "max_dX": dX,
                                                      "iteration_difference_max": int(np.abs(di).max()), "same_iterations_frac": float((di == 0).mean())})
    # measured: bench 2e-9, script60 6e-11; tracking40 (hard, barely feasible instances) 2.8e-6 in the thrusts, 4e-8 in the states
    assert dX < TOL and dS < 1e-4 and dT < (1e-5 if scenario == "tracking40_qp8" else TOL), (dT, dS, dX)


def test_diagnostics_agree_with_the_oracle_evaluation(cuda_device):
    """bench.py reports solver.max_explicit_res_* through mpc_blaster_b200.diagnostics (torch, on the device); it must
    compute what the NumPy oracle computes from the same export."""
    from mpc_blaster_b200 import diagnostics
    N, B = 20, 128
    P = bo.canonical_problem(N)
    x0, yref = sc.random_setpoints(B, seed=8)
    mpc = _mpc(N, B)
    mpc.reset(x0, sc.hover_trim())
    mpc.solve(x0, yref)
    r = _kkt(mpc, P, B)
    d = diagnostics.explicit_kkt_residuals(mpc, B)
    for k in ("stat", "eq", "ineq", "comp"):
        a, b = d[k].cpu().numpy(), r[k]
        assert np.allclose(a, b, rtol=1e-7, atol=1e-11), (k, np.abs(a - b).max())  # summation order differs (terms are O(1e3))


# --------------------------------------------------------------------------- full-size parity, configs 3 / 5 / 4
def test_full_size_config3_tracking_sample_against_oracle(cuda_device):
    """BASELINE config 3 at its full size: 65,536 instances, N = 40, per-stage yref on the lemniscate (input and state
    bounds active), default scheduler (workspace chunks on the persistent four-instances-per-warp kernel).  A random
    sample of 512 instances spread over all chunks against the C oracle: status, IPM iteration count, the whole
    trajectory X, U -- plus the whole-batch convergence rate."""
    B, N = 65_536, 40
    P = bo.canonical_problem(N)
    x0, yref = sc.lemniscate_tracking(B, N)
    trim = sc.hover_trim()
    mpc = _mpc(N, B)
    mpc.reset(x0, trim)
    yref_d = torch.as_tensor(yref, device="cuda")
    u0, X, U, st = mpc.solve(x0, yref_d)
    conv = float((st == 0).double().mean())
    assert conv > 0.99, conv
    idx = np.sort(np.random.default_rng(3).choice(B, 512, replace=False))
    orc = co.BatchRTI(P, len(idx))
    orc.reset(x0[idx], trim)
    uo, Xo, Uo, sto = orc.solve(x0[idx], yref[idx])
    ti = torch.as_tensor(idx, device="cuda")
    assert (st[ti].cpu().numpy() == sto).all()
    # the tracking batch contains hard instances (up to ~35 interior-point iterations, barely feasible): there the two
    # implementations (Householder LQ on 8 lanes per instance against the oracle's scalar LQ) may stop an iteration
    # apart; everywhere else the counts must be equal, and the converged points agree at 1e-6 either way
    di = mpc.iters[ti].cpu().numpy() - orc.iters
    same_frac = float((di == 0).mean())
    assert same_frac > 0.97 and np.abs(di).max() <= 2, (same_frac, di.min(), di.max())
    ok = (sto == 0) & (di == 0)
    late = (sto == 0) & (di != 0)   # stopped an iteration apart: the same point to the looser bound
    if late.any():
        assert np.abs(U[ti].cpu().numpy()[late] - Uo[late]).max() < 1e-4 and np.abs(X[ti].cpu().numpy()[late] - Xo[late]).max() < 1e-5
    dUi = np.abs(U[ti].cpu().numpy()[ok] - Uo[ok]).reshape(int(ok.sum()), -1).max(1)
    dU, dU99 = dUi.max(), np.percentile(dUi, 99)
    dX = np.abs(X[ti].cpu().numpy()[ok] - Xo[ok]).max()
    # active sets of the sample ("report active-set sizes")
    act_u = int(((np.abs(Uo[ok] - P.lbu) < 1e-6) | (np.abs(Uo[ok] - P.ubu) < 1e-6)).sum(axis=(1, 2)).mean())
    act_x = int(((np.abs(Xo[ok][:, 1:N] - P.lbx) < 1e-6) | (np.abs(Xo[ok][:, 1:N] - P.ubx) < 1e-6)).sum(axis=(1, 2)).mean())
    _report("r02_config3_full.json", {"B": B, "N": N, "converged_frac": conv, "sample": len(idx), "max_dU": float(dU), "p99_dU": float(dU99), "max_dX": float(dX), "same_ipm_iteration_count_frac": same_frac,
                                      "mean_active_input_bounds": act_u, "mean_active_state_bounds": act_x,
                                      "mean_ipm_iters": float(mpc.iters.double().mean())})
    # measured: states 2e-8; inputs 1.4e-6 on the worst of 512 sampled instances (a thrust of a barely feasible instance: the
    # unrefined solves of both sides are accurate to a few 1e-6 there, see the strict-vs-default test), 99 % within 1e-6
    assert dX < TOL and dU99 < TOL and dU < 1e-5, (dU, dU99, dX)
    assert act_u > 10 and act_x > 5  # the scenario does what config 3 asks for: both kinds of bounds bind
    del mpc
    torch.cuda.empty_cache()


@pytest.mark.parametrize("N", [40, 80])
def test_full_size_config5_long_horizons_sample_against_oracle(cuda_device, N):
    """BASELINE config 5 at its full size for the long horizons: 1,048,576 BLASTER17 instances, N = 40 and N = 80 (the
    N = 20 column is test_full_size_config5_sample_against_oracle), workspace walked in chunks.  512 sampled instances
    against the C oracle: status (N = 80 leaves ~3.7 % of the linearised QPs infeasible: those must report the
    oracle's status after the oracle's number of iterations), iteration count and u0."""
    B = 1 << 20
    P = bo.canonical_problem(N)
    x0, yref = sc.random_setpoints(B, seed=4567)
    trim = sc.hover_trim()
    mpc = _mpc(N, B)
    mpc.reset(x0, trim)
    u0, _, _, st = mpc.solve(x0, yref, want_traj=False)
    conv = float((st == 0).double().mean())
    idx = np.sort(np.random.default_rng(N).choice(B, 512, replace=False))
    orc = co.BatchRTI(P, len(idx))
    orc.reset(x0[idx], trim)
    uo, _, _, sto = orc.solve(x0[idx], yref[idx])
    ti = torch.as_tensor(idx, device="cuda")
    assert (st[ti].cpu().numpy() == sto).all() and (mpc.iters[ti].cpu().numpy() == orc.iters).all()
    ok = sto == 0
    du = float(np.abs(u0[ti].cpu().numpy()[ok] - uo[ok]).max())
    _report(f"r02_config5_full_N{N}.json", {"B": B, "N": N, "converged_frac": conv, "sample_converged_frac": float(ok.mean()), "max_du0": du})
    assert du < TOL and conv > (0.97 if N == 40 else 0.94), (du, conv)
    del mpc
    torch.cuda.empty_cache()


def test_full_size_config4_closed_loop_with_lockstep_oracle_subset(cuda_device):
    """BASELINE config 4 at its full size: 16,384 vehicles x 500 closed-loop control steps entirely on the device
    (mpcb_closed_loop: solve, plant step, bookkeeping; un-shifted warm start).  The first 64 vehicles are also stepped
    one control step at a time on a second handle with the C oracle in lock-step: before every step the oracle is given
    the GPU's iterate and state, both solve, and status, iteration count, u0 and the new iterate must agree at 1e-6 for
    all 500 steps.  The step-wise 64 must end bit-identical to the first 64 of the 16,384 device loop.  (Two
    free-running loops drift apart chaotically in the swivel-rate components -- DESIGN.md section 2 -- so the free-running
    oracle's drift is reported, not asserted.)"""
    B, N, steps, S = 16_384, 20, 500, 64
    P = bo.canonical_problem(N)
    x0, yref = sc.closed_loop_setpoints(B, seed=3456)
    trim = sc.hover_trim()
    big = _mpc(N, B)
    big.reset(x0, trim)
    xf, ul, nfail, its = big.closed_loop(x0, yref, steps=steps)
    torch.cuda.synchronize()
    fail_frac = float((nfail > 0).double().mean())
    small = _mpc(N, S, qp8_batch=1)  # the 16,384-instance loop runs on the four-instances-per-warp kernel: same kernel here
    small.reset(x0[:S], trim)
    orc = co.BatchRTI(P, S)
    x = torch.as_tensor(x0[:S], device="cuda")
    yr = torch.as_tensor(yref[:S], device="cuda")
    worst_u = worst_x = 0.0
    n_bad = 0
    for t in range(steps):
        Xi, Ui = small.iterate(S)
        orc.X[:], orc.U[:] = Xi.cpu().numpy(), Ui.cpu().numpy()
        xh = x.cpu().numpy()
        u0, X, U, st = small.solve(x, yr)
        uo, Xo, Uo, sto = orc.solve(xh, yref[:S])
        stg = st.cpu().numpy()
        assert (stg == sto).all(), (t, stg, sto)
        ok = sto == 0
        n_bad += int((~ok).sum())
        assert (small.iters.cpu().numpy()[ok] == orc.iters[ok]).all(), t
        if ok.any():
            worst_u = max(worst_u, float(np.abs(U.cpu().numpy()[ok] - Uo[ok]).max()))
            worst_x = max(worst_x, float(np.abs(X.cpu().numpy()[ok] - Xo[ok]).max()))
        x = small.step_plant(x, u0)
    assert worst_u < TOL and worst_x < TOL, (worst_u, worst_x)
    assert torch.equal(x, xf[:S]) and torch.equal(u0, ul[:S])
    # free-running oracle on the same 64 vehicles: reported drift
    free = co.BatchRTI(P, S)
    free.reset(x0[:S], trim)
    xo = x0[:S].copy()
    for t in range(steps):
        uo, _, _, _ = free.solve(xo, yref[:S])
        xo = co.plant_step(P, xo, uo)
    drift = np.abs(xo - x.cpu().numpy())
    _report("r02_config4_full.json", {"B": B, "N": N, "steps": steps, "instances_with_a_failed_step_frac": fail_frac,
                                      "mean_ipm_iters_per_step": float(its.double().mean()) / steps,
                                      "lockstep_subset": S, "lockstep_max_dU": worst_u, "lockstep_max_dX": worst_x,
                                      "lockstep_failed_solves": n_bad,
                                      "free_running_drift_after_500_steps": {"position_max": float(drift[:, 0:3].max()),
                                                                             "all_states_max": float(drift.max()),
                                                                             "position_median": float(np.median(drift[:, 0:3].max(1)))},
                                      "final_position_error_median": float(np.median(np.abs(xf.cpu().numpy()[:, 0:3] - yref[:, 0:3]).max(1)))})
    assert fail_frac < 0.05, fail_frac
    del big, small
    torch.cuda.empty_cache()


# --------------------------------------------------------------------------- SQP to convergence, strict semantics
def test_sqp_to_convergence_matches_c_oracle(cuda_device):
    """SURVEY 8f row 1 with the options of the reference's dump (nlp_solver_tol_* = 1e-6, nlp_solver_max_iter = 100,
    acados_ocp_blasterModel.json solver_options): per-instance NLP residuals on the device, finished instances skipped
    by the kernels, against the C oracle running the same loop: SQP status, number of QPs, total interior-point
    iterations, residuals, and ALL of X, U at north_star's 1e-6."""
    B, N = 96, 20
    P = bo.canonical_problem(N)
    x0, yref = sc.random_setpoints(B, seed=77)
    trim = sc.hover_trim()
    mpc = _mpc(N, B)
    # the QPs of this mode are solved with the reference-semantics rule set (refined multipliers), include/mpcb.h
    orc = co.BatchRTI(P, B, strict=True, max_iter=int(mpc.cfg.ipm_max_iter))
    mpc.reset(x0, trim)
    orc.reset(x0, trim)
    u0, X, U, st = mpc.solve(x0, yref, sqp_iters=100, sqp_tol=1e-6)
    uo, Xo, Uo, sto, n_qp, qp_it, res = orc.sqp_solve(x0, yref, max_iter=100, tol=1e-6)
    st = st.cpu().numpy()
    assert (st == sto).all(), (st, sto)
    conv = sto == 0
    assert conv.mean() > 0.9, conv.mean()
    assert (mpc.sqp_iters.cpu().numpy() == n_qp).all() and (mpc.iters.cpu().numpy() == qp_it).all()
    assert (n_qp[conv] >= 2).all() and len(set(n_qp[conv].tolist())) > 1  # instances do finish at different iterations
    gres = mpc.nlp_res.cpu().numpy()
    assert (gres[conv] <= 1e-6).all() and np.abs(gres[conv] - res[conv]).max() < 1e-7
    assert np.abs(X.cpu().numpy()[conv] - Xo[conv]).max() < TOL and np.abs(U.cpu().numpy()[conv] - Uo[conv]).max() < TOL
    assert np.abs(u0.cpu().numpy()[conv] - uo[conv]).max() < TOL
    # the converged iterate is a fixed point: a second call (multipliers restart at zero [upstream D4], so the first residual
    # evaluation cannot pass where bounds are active) needs at most one QP and stays where it is
    u1, X1, U1, st1 = mpc.solve(x0, yref, sqp_iters=100, sqp_tol=1e-6)
    tc = torch.as_tensor(conv, device="cuda")
    n2 = mpc.sqp_iters.cpu().numpy()[conv]
    assert (n2 <= 3).all() and np.median(n2) <= 1 and (st1.cpu().numpy()[conv] == 0).all(), (n2.max(), np.median(n2))
    assert float((X1[tc] - X[tc]).abs().max()) < 1e-7 and float((U1[tc][..., :4] - U[tc][..., :4]).abs().max()) < TOL
    # iteration cap: status 2 after exactly that many QPs
    mpc.reset(x0, trim)
    _, _, _, st2 = mpc.solve(x0, yref, sqp_iters=2, sqp_tol=1e-6)
    assert (st2.cpu().numpy() == 2).all() and (mpc.sqp_iters.cpu().numpy() == 2).all()
    _report("r02_sqp.json", {"B": B, "N": N, "converged_frac": float(conv.mean()), "sqp_iters_mean": float(n_qp[conv].mean()),
                             "sqp_iters_max": int(n_qp[conv].max()), "max_dX": float(np.abs(X.cpu().numpy()[conv] - Xo[conv]).max()),
                             "max_dU": float(np.abs(U.cpu().numpy()[conv] - Uo[conv]).max()), "max_nlp_res": float(gres[conv].max())})


def test_strict_reference_semantics_match_c_oracle(cuda_device):
    """mpcb_config.strict_reference (explicit residual norms in the stopping test, no divergence exit, cap 500, last
    iterate applied on max-iter) on the GPU against the C oracle with the same switch: status, iteration count, iterate.
    The config-1 closed loop (zero iterate, state bounds binding from the first step) is stepped 25 times with the
    oracle re-seeded from the GPU's iterate; the number of steps whose explicit stationarity norm never reaches 1e-6
     is reported."""
    N, B = 20, 96
    P = bo.canonical_problem(N)
    x0, yref = sc.random_setpoints(B, seed=21)
    trim = sc.hover_trim()
    mpc = _mpc(N, B, strict_reference=True)
    assert mpc.cfg.ipm_max_iter == 500
    orc = co.BatchRTI(P, B, strict=True)
    mpc.reset(x0, trim)
    orc.reset(x0, trim)
    u0, X, U, st = mpc.solve(x0, yref)
    uo, Xo, Uo, sto = orc.solve(x0, yref)
    assert (st.cpu().numpy() == sto).all() and (mpc.iters.cpu().numpy() == orc.iters).all()
    take = (sto == 0) | (sto == 2)
    assert (sto == 0).mean() > 0.8
    assert np.abs(X.cpu().numpy()[take] - Xo[take]).max() < TOL and np.abs(U.cpu().numpy()[take] - Uo[take]).max() < TOL
    # config 1 closed loop under the reference's semantics
    one = _mpc(N, 1, strict_reference=True)
    o1 = co.BatchRTI(P, 1, nthreads=1, strict=True)
    xs, yr = bo.canonical_x0_yref()
    x = xs.reshape(1, 17).copy()
    hist = []
    for step in range(25):
        Xi, Ui = one.iterate(1)
        o1.X[:], o1.U[:] = Xi.cpu().numpy(), Ui.cpu().numpy()
        u0, X, U, st = one.solve(x, yr.reshape(1, -1))
        uo, Xo, Uo, sto = o1.solve(x, yr.reshape(1, -1))
        assert int(st[0]) == int(sto[0]) and int(one.iters[0]) == int(o1.iters[0]), (step, int(st[0]), int(sto[0]), int(one.iters[0]), int(o1.iters[0]))
        assert int(sto[0]) in (0, 2)
        assert np.abs(U.cpu().numpy() - Uo).max() < TOL and np.abs(X.cpu().numpy() - Xo).max() < TOL
        hist.append((int(sto[0]), int(o1.iters[0])))
        x = co.plant_step(P, x, uo)
    _report("r02_strict.json", {"batch_converged_frac": float((sto == 0).mean()), "config1_steps": len(hist),
                                "config1_steps_hitting_the_500_cap": sum(1 for s, _ in hist if s == 2),
                                "config1_iters": [i for _, i in hist]})


def test_from_acados_json_takes_the_iteration_cap_of_the_dump(cuda_device):
    """qp_solver_iter_max of the reference's dump (500, acados_ocp_blasterModel.json solver_options) reaches the
    solver unless the caller overrides it."""
    from mpc_blaster_b200 import BlasterMPC
    path = os.path.join(ROOT, "tests", "golden", "acados_ocp_subset.json")
    a = BlasterMPC.from_acados_json(path, N=20, batch=4)
    assert a.cfg.ipm_max_iter == 500
    b = BlasterMPC.from_acados_json(path, N=20, batch=4, ipm_max_iter=60)
    assert b.cfg.ipm_max_iter == 60


def test_host_entry_point_validates_shapes_and_restores_the_device(cuda_device):
    """solve_host checks the shapes of yref / p before the library copies B*ny / B*N*25 doubles from the pointers, and the
    entry points that make the handle's GPU current put the caller's device back."""
    from mpc_blaster_b200 import BlasterMPC
    N, B = 6, 5
    mpc = BlasterMPC.canonical(N=N, batch=B, device="cuda")  # no index: the current CUDA device
    assert mpc.device.index == torch.cuda.current_device()
    x0, yref = sc.random_setpoints(B, seed=1)
    mpc.solve_host(x0, yref)
    with pytest.raises(ValueError):
        mpc.solve_host(x0, np.zeros((N + 1, 23)))         # [N+1, ny] is not [B, ny]
    with pytest.raises(ValueError):
        mpc.solve_host(x0, yref, p=np.zeros(25 * B))      # flattened p
    with pytest.raises(ValueError):
        mpc.solve_host(x0, yref.reshape(1, B, 23, 1))     # ndim 4
    if torch.cuda.device_count() >= 2:
        torch.cuda.set_device(0)
        other = BlasterMPC.canonical(N=N, batch=B, device="cuda:1")
        assert torch.cuda.current_device() == 0
        u0, _, _, st = other.solve_host(x0, yref)
        assert torch.cuda.current_device() == 0 and (st == 0).all()
        t = torch.zeros(4, device="cuda")
        assert t.device.index == 0
        del other
        assert torch.cuda.current_device() == 0


# --------------------------------------------------------------------------- sharded == unsharded on GPUs
def test_sharded_solve_equals_unsharded_on_gpus(cuda_device, tmp_path):
    """SURVEY section 4, distributed row: a global batch solved as contiguous shards by G ranks (one process per GPU,
    NCCL all-gather of u0 / status) is bit-identical to the same batch solved on one GPU.  With a single visible GPU the
    ranks share it and gather over gloo -- the sharding and the kernels are the same."""
    G = min(torch.cuda.device_count(), 4)
    ranks = G if G >= 2 else 2
    out = tmp_path / "ok.json"
    env = dict(os.environ, MPCB_IDENTITY_OUT=str(out))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tools", "sharded_identity.py"), "--batch", "1000"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rep = json.load(open(out))
    assert rep["identical"] and rep["world"] == ranks and rep["converged_frac"] > 0.95
    _report("r02_sharded_identity.json", rep)

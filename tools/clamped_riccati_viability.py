"""Experiment (test infrastructure, not product): is north_star's "clamped Riccati with a projected line
search" a faster way to the box-constrained QP than the interior point, where it can apply at all?

It can only apply when no state bound is active, i.e. BASELINE config 3b: figure-eight
tracking, N = 40, state bounds widened x1000, 151 of 240 input bounds active at the solution.  The method is
projected Newton on the QP in the inputs (Bertsekas 1982, what control-limited DDP does stage by stage): free /
clamped split from the sign of the gradient at the bounds, Newton step on the free inputs -- one Riccati
factorisation with the clamped inputs removed --, then a backtracking line search along the projection
(clamping) of that step onto the box.  Here the Newton systems are solved densely on the condensed problem,
which counts factorisations exactly as a Riccati implementation would need them; the interior point is the
checker's (oracle/blaster_oracle.py), whose iteration count is the CUDA path's.

    python tools/clamped_riccati_viability.py        # output recorded in profiles/r01_clamped_riccati_viability.txt
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_blaster_b200 import scenarios as sc  # noqa: E402
from oracle import blaster_oracle as bo  # noqa: E402


def condensed(qp):
    """Eliminate the states: J(u) = 1/2 u'Mu + r'u over the stacked input increments."""
    H, g, C, c, lb, ub = bo.qp_to_dense(qp)
    N, nx, nu = qp.A.shape[0], qp.A.shape[1], qp.B.shape[2]
    nz = nu + nx
    iu = np.concatenate([np.arange(k * nz, k * nz + nu) for k in range(N)])
    ix = np.setdiff1d(np.arange(N * nz), iu)
    X0 = np.linalg.solve(C[:, ix], c)
    Xu = -np.linalg.solve(C[:, ix], C[:, iu])
    M = np.diag(H[iu]) + Xu.T @ (H[ix][:, None] * Xu)
    r = g[iu] + Xu.T @ (H[ix] * X0 + g[ix])
    return M, r, lb[iu], ub[iu], iu


EPS_ACT = float(os.environ.get("EPS_ACT", "1e-2"))  # fraction of the box width


def projected_newton(M, r, lb, ub, tol=1e-6, max_fact=200):
    u = np.clip(np.zeros_like(r), lb, ub)
    J = lambda v: 0.5 * v @ M @ v + r @ v  # noqa: E731
    for nfact in range(max_fact):
        gr = M @ u + r
        # Bertsekas' epsilon-active set: bounds within eps_k whose gradient points outwards count as clamped
        eps_k = min(EPS_ACT, np.linalg.norm(u - np.clip(u - gr, lb, ub)))
        act = ((u <= lb + eps_k * (ub - lb)) & (gr > 0)) | ((u >= ub - eps_k * (ub - lb)) & (gr < 0))
        at_bound = ((u <= lb + 1e-10) & (gr > 0)) | ((u >= ub - 1e-10) & (gr < 0))
        if np.abs(np.where(at_bound, 0.0, gr)).max() <= tol:
            return u, nfact, True
        f = ~act
        d = np.zeros_like(u)
        d[f] = -np.linalg.solve(M[np.ix_(f, f)], gr[f])     # = one clamped Riccati factorisation + solve
        a, J0 = 1.0, J(u)
        while True:                                          # projected (clamped) backtracking line search
            un = np.clip(u + a * d, lb, ub)
            if J(un) <= J0 + 1e-4 * gr @ (un - u) or a < 1e-12:
                break
            a *= 0.5
        u = un
    return u, max_fact, False


def main():
    N, B = 40, 32
    P = bo.canonical_problem(N)
    P.lbx, P.ubx = np.full(17, -1e3), np.full(17, 1e3)      # config 3b: only input bounds can bind
    x0s, yrefs = sc.lemniscate_tracking(B, N)
    nf, ok, ipm, nact, err = [], [], [], [], []
    for i in range(B):
        X = np.repeat(x0s[i][None], N + 1, 0)
        U = np.tile(sc.hover_trim(), (N, 1))
        qp = bo.build_qp(X, U, x0s[i], yrefs[i], None, P)
        M, r, lb, ub, iu = condensed(qp)
        u, nfact, conv = projected_newton(M, r, lb, ub)
        H, g, C, c, lbz, ubz = bo.qp_to_dense(qp)
        res = bo.ipm_dense(H, g, C, c, lbz, ubz)
        nf.append(nfact); ok.append(conv); ipm.append(res.iters)
        nact.append(int(((u <= lb + 1e-9) | (u >= ub - 1e-9)).sum()))
        if conv:
            err.append(np.abs((u - res.z[iu]).reshape(N, 6)[:, :4]).max())
    print(f"config 3b flavour: {B} instances, N = {N}, inputs only bounded; {np.mean(nact):.0f} of {6 * N} input bounds active on average")
    print(f"interior point (the shipped algorithm): {np.mean(ipm):.1f} factorisations on average, {np.max(ipm)} at worst, all converged")
    print(f"clamped Riccati + projected line search: {np.mean(nf):.1f} factorisations on average, {np.max(nf)} at worst, "
          f"{int(np.sum(ok))} of {B} converged within 200 (projected gradient <= 1e-6)")
    if err:
        print(f"   where it converged: max |du - du_ipm| over the thrusts = {np.max(err):.2e}")


if __name__ == "__main__":
    main()

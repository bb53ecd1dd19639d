#!/usr/bin/env python
"""Mixed-precision viability, CPU part: the C restatement of the kernels' iteration with the
stage factorisation (W = [B A]'L, Gram matrix, Cholesky) done in FP32 while `mixed_mu` < mu <= mu0 -- i.e. only in the
iterations where the product already accepts an inexact Newton direction (it runs the normal-equations form there) -- and
everything else (residuals, right-hand sides, substitutions, the FP64 Householder LQ of the tail) in FP64.

Reports, against the all-FP64 iteration on the same inputs: converged fraction, change of the interior-point iteration
count, and the final |du| (thrusts / swivel rates) and |dx|.  Usage: python tools/mixed_precision_viability.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_blaster_b200 import scenarios as sc  # noqa: E402
from oracle import blaster_oracle as bo, c_oracle as co  # noqa: E402


def run(name, P, x0, yref, B):
    trim = sc.hover_trim(P.nu)
    ref = co.BatchRTI(P, B)
    ref.reset(x0, trim)
    _, Xr, Ur, sr = ref.solve(x0, yref)
    print(f"== {name}: B={B} N={P.N}  FP64: converged {np.mean(sr == 0):.4f}, iterations mean {ref.iters.mean():.3f} max {ref.iters.max()}")
    for thr in (1e-1, 1e-2, 1e-3, 1e-4, 1e-5):
        m = co.BatchRTI(P, B, mixed_mu=thr)
        m.reset(x0, trim)
        _, X, U, s = m.solve(x0, yref)
        both = (s == 0) & (sr == 0)
        di = m.iters[both] - ref.iters[both]
        nu4 = min(4, P.nu)
        if not both.any():
            print(f"  FP32 factorisation while mu > {thr:g}: converged {np.mean(s == 0):.4f} -- nothing left to compare")
            continue
        dt = np.abs(U[both][..., :nu4] - Ur[both][..., :nu4]).max()
        ds = np.abs(U[both][..., nu4:] - Ur[both][..., nu4:]).max() if P.nu > 4 else 0.0
        dx = np.abs(X[both] - Xr[both]).max()
        print(f"  FP32 factorisation while mu > {thr:g}: converged {np.mean(s == 0):.4f} (FP64-converged lost: {int(((sr == 0) & (s != 0)).sum())}), "
              f"iterations mean {m.iters.mean():.3f} (changed on {np.mean(di != 0) * 100:.1f} %, max +{di.max()} / {di.min()}), "
              f"max|d thrust| {dt:.2e}  |d swivel| {ds:.2e}  |dx| {dx:.2e}")


if __name__ == "__main__":
    form = "Householder LQ in FP32" if os.environ.get("ORC_MIXED_LQ") else "Gram matrix + Cholesky in FP32"
    print(f"### early-iteration stage factorisation: {form} (status 4 = a pivot is not positive)")
    P = bo.canonical_problem(20)
    x0, yref = sc.random_setpoints(1024, seed=1234)
    run("configs[1] bench batch", P, x0, yref, 1024)
    P40 = bo.canonical_problem(40)
    x0, yref = sc.lemniscate_tracking(512, 40)
    run("config 3 tracking (state bounds active)", P40, x0, yref, 512)
    P12 = bo.canonical_problem(20, 12)
    x0, yref = sc.random_setpoints(1024, seed=1234, nx=12, nu=4)
    run("QUAD12 bench batch", P12, x0, yref, 1024)

#!/usr/bin/env python
"""Numerics of the split P-form stage factorisation (C restatement, ric_alg = 3) against the Householder LQ (ric_alg = 1,
the checker): statuses, interior-point iteration counts and solutions on the bench batch, the N = 40 tracking batch and
QUAD12, two RTI steps each; and how long the list of carried huge columns gets.

    ORC_HUGE_TAU=1e6 ORC_HUGE_TAU_D=1e6 python tools/split_factor_viability.py
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_blaster_b200 import scenarios as sc  # noqa: E402
from oracle import blaster_oracle as bo  # noqa: E402
from oracle import c_oracle as co  # noqa: E402


def run(name, P, x0, yref, steps=2, init=True):
    B = x0.shape[0]
    a = co.BatchRTI(P, B, ric_alg=1)
    b = co.BatchRTI(P, B, ric_alg=3)
    if init:
        a.reset(x0, sc.hover_trim(P.nu))
        b.reset(x0, sc.hover_trim(P.nu))
    hist = (C.c_long * 36)()
    co.lib().orc_huge_histogram(hist, 1)
    co.lib().orc_split_iterations((C.c_long * 2)(), 1)
    xa = x0.copy()
    for s in range(steps):
        X_pre, U_pre = a.X.copy(), a.U.copy()
        ua, Xa, Ua, sa = a.solve(xa, yref)
        ia = a.iters.copy()
        # the experiment starts every step from the checker's iterate: single-solve differences, not a drifting loop
        b.X[:], b.U[:] = X_pre, U_pre
        ub, Xb, Ub, sb = b.solve(xa, yref)
        ib = b.iters.copy()
        ok = (sa == 0) & (sb == 0)
        dU = np.abs(Ua - Ub)[ok]
        dX = np.abs(Xa - Xb)[ok]
        print(f"{name:28s} step {s}: status equal {np.mean(sa == sb):.4f}  ok {ok.mean():.4f}  iterations equal {np.mean(ia[ok] == ib[ok]):.4f} "
              f"(max diff {np.abs(ia[ok] - ib[ok]).max()}; mean {ia[ok].mean():.2f} / {ib[ok].mean():.2f})  max|dU| thrust {dU[..., :4].max():.2e} other {dU[..., 4:].max() if P.nu > 4 else 0:.2e}  max|dX| {dX.max():.2e}",
              flush=True)
        xa = co.plant_step(P, xa, ua)
    co.lib().orc_huge_histogram(hist, 1)
    si = (C.c_long * 2)()
    co.lib().orc_split_iterations(si, 1)
    print(f"   interior-point iterations factorised by the split recursion: {100 * si[1] / max(si[0] + si[1], 1):.1f} %")
    h = np.array(list(hist), dtype=np.int64)
    tot = max(h.sum(), 1)
    print("   list length per stage factorisation: " + "  ".join(f"{i}:{100 * v / tot:.1f}%" for i, v in enumerate(h) if v), flush=True)


def main():
    P = bo.canonical_problem(20, 17)
    x0, yref = sc.random_setpoints(1024, seed=1234)
    run("bench batch 1024 x N=20", P, x0, yref)
    P40 = bo.canonical_problem(40, 17)
    x0, yref = sc.lemniscate_tracking(256, 40)
    run("tracking 256 x N=40", P40, x0, yref)
    x0, yref = sc.random_setpoints(256, seed=77)
    run("random 256 x N=40 zero iterate", P40, x0, yref, init=False)
    P12 = bo.canonical_problem(20, 12)
    x0, yref = sc.random_setpoints(512, seed=1234, nx=12, nu=4)
    run("QUAD12 512 x N=20", P12, x0, yref)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""BASELINE.json config 4: closed-loop Monte-Carlo, B quadrotors x S control steps with
un-shifted warm-started SQP-RTI, entirely on the device (mpcb_closed_loop).  Reports
instance-steps/s and how many solves failed; optionally checks a subset against the C oracle
run in lock-step (iterate re-seeded from the GPU each step, see DESIGN.md on parity)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_blaster_b200 import BlasterMPC, scenarios as sc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16384)
ap.add_argument("--steps", type=int, default=500)
ap.add_argument("--horizon", type=int, default=20)
ap.add_argument("--check", type=int, default=0, help="instances to check against the C oracle (first steps only)")
ap.add_argument("--max-fail-frac", type=float, default=1e-4,
                help="fail if more than this fraction of all solves reports a non-zero status (a failed solve applies the stale u0, include/mpcb.h)")
a = ap.parse_args()

B, S, N = a.batch, a.steps, a.horizon
x0, yref = sc.closed_loop_setpoints(B, seed=3456)
mpc = BlasterMPC.canonical(N=N, batch=B)
mpc.reset(x0, sc.hover_trim())
torch.cuda.synchronize()
t = time.perf_counter()
xT, u_last, n_fail, iters = mpc.closed_loop(x0, yref, steps=S)
torch.cuda.synchronize()
dt = time.perf_counter() - t
err = float((xT[:, :3] - torch.as_tensor(yref[:, :3], device="cuda")).norm(dim=1).mean())
out = dict(config="closed-loop Monte-Carlo", batch=B, steps=S, horizon=N, seconds=dt, instance_steps_per_s=B * S / dt,
           failed_solves=int(n_fail.sum()), instances_with_a_failed_solve=int((n_fail > 0).sum()),
           mean_ipm_iters=float(iters.double().mean() / S), mean_final_position_error_m=err,
           kernel_launches=mpc.kernel_launches())
print(json.dumps(out), flush=True)
assert out["failed_solves"] <= a.max_fail_frac * B * S, f"{out['failed_solves']} failed solves of {B * S}"

if a.check:
    from oracle import blaster_oracle as bo, c_oracle as co
    C = a.check
    P = bo.canonical_problem(N)
    g = BlasterMPC.canonical(N=N, batch=C)
    g.reset(x0[:C], sc.hover_trim())
    orc = co.BatchRTI(P, C)
    x = x0[:C].copy()
    worst = 0.0
    for s in range(10):
        Xg, Ug = g.iterate()
        orc.X[:], orc.U[:] = Xg.cpu().numpy(), Ug.cpu().numpy()
        u0, X, U, st = g.solve(x, yref[:C])
        uo, Xo, Uo, sto = orc.solve(x, yref[:C])
        assert (st.cpu().numpy() == sto).all(), (s, st.cpu().numpy(), sto)
        ok = sto == 0
        worst = max(worst, float(np.abs(U.cpu().numpy()[ok] - Uo[ok]).max()), float(np.abs(X.cpu().numpy()[ok] - Xo[ok]).max()))
        x = g.step_plant(x, u0).cpu().numpy()
    print(json.dumps(dict(check_instances=C, check_steps=10, max_abs_diff_vs_oracle=worst)), flush=True)

#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): 8 instances, N=6, both
model variants, one solve each plus plant step, cost and command map; then the four-instances-per-warp kernel with
refilled groups."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_blaster_b200 import BlasterMPC, scenarios as sc  # noqa: E402

for variant in (17, 12):
    nx, nu = (17, 6) if variant == 17 else (12, 4)
    x0, yref = sc.random_setpoints(8, seed=3, nx=nx, nu=nu)
    mpc = BlasterMPC.canonical(N=6, batch=8, variant=variant)
    mpc.reset(x0, sc.hover_trim(nu))
    u0, X, U, st = mpc.solve(x0, yref)
    xn = mpc.step_plant(x0, u0)
    c = mpc.cost(yref)
    q, t = mpc.command_map(X[:, 0], u0) if variant == 17 else (None, None)
    torch.cuda.synchronize()
    print(variant, st.tolist(), mpc.iters.tolist(), float(c.sum()))
# four-instances-per-warp kernel, one persistent warp: its groups are refilled from the work counter
for variant in (17, 12):
    nx, nu = (17, 6) if variant == 17 else (12, 4)
    x0, yref = sc.random_setpoints(11, seed=5, nx=nx, nu=nu)
    mpc = BlasterMPC.canonical(N=5, batch=11, variant=variant, qp8_batch=1, qp8_warps=1)
    mpc.reset(x0, sc.hover_trim(nu))
    u0, X, U, st = mpc.solve(x0, yref)
    torch.cuda.synchronize()
    print("qp8", variant, st.tolist(), mpc.iters.tolist())
print("done")
# round 2: reference-semantics instantiation (explicit norms + iterative refinement), SQP to convergence (NLP residual and
# bookkeeping kernels), and the debug export of the QP / interior-point iterate
for variant in (17, 12):
    nx, nu = (17, 6) if variant == 17 else (12, 4)
    x0, yref = sc.random_setpoints(6, seed=9, nx=nx, nu=nu)
    mpc = BlasterMPC.canonical(N=6, batch=6, variant=variant, strict_reference=True)
    mpc.reset(x0, sc.hover_trim(nu))
    u0, X, U, st = mpc.solve(x0, yref)
    d = mpc.debug_qp()
    torch.cuda.synchronize()
    print("strict", variant, st.tolist(), mpc.iters.tolist(), float(d["z"].abs().max()))
    mpc2 = BlasterMPC.canonical(N=6, batch=6, variant=variant)
    mpc2.reset(x0, sc.hover_trim(nu))
    u0, X, U, st = mpc2.solve(x0, yref, sqp_iters=20, sqp_tol=1e-6)
    torch.cuda.synchronize()
    print("sqp", variant, st.tolist(), mpc2.sqp_iters.tolist(), mpc2.nlp_res.max(dim=0).values.tolist())
print("done round 2")

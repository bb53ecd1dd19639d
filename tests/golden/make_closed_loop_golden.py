"""Golden closed-loop trajectory of BASELINE.json config 1 (single BLASTER quadrotor, hover to
set-point, N=20: reference simulation_blaster.py:47-48 and its loop :56-105) produced by OUR C
oracle -- the reference ships no recorded trajectory and its solver cannot run here, so this file
pins the oracle against regressions, not against the reference.  The un-shifted RTI iterate
before each of the first LOCK control steps is stored too, so the GPU path can be checked in
lock-step (free-running loops drift apart in the weakly determined swivel-rate directions, see
DESIGN.md "what 1e-6 parity means").

    python tests/golden/make_closed_loop_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import blaster_oracle as bo  # noqa: E402
from oracle import c_oracle as co  # noqa: E402

STEPS = 200
LOCK = 60
P = bo.canonical_problem(20)
x0, yref = bo.canonical_x0_yref()
c = co.BatchRTI(P, 1, nthreads=1)
x = x0[None].copy()
simX, simU, iters, itX, itU = [x[0].copy()], [], [], [], []
for s in range(STEPS):
    if s <= LOCK:
        itX.append(c.X[0].copy())
        itU.append(c.U[0].copy())
    u0, X, U, st = c.solve(x, yref)
    assert st[0] == 0
    iters.append(int(c.iters[0]))
    x = co.plant_step(P, x, u0)
    simU.append(u0[0].copy())
    simX.append(x[0].copy())
np.savez(os.path.join(HERE, "hover_closed_loop_golden.npz"), simX=np.array(simX), simU=np.array(simU), iters=np.array(iters),
         yref=yref, N=20, itX=np.array(itX), itU=np.array(itU))
print("final z %.4f, mean IPM iterations %.2f" % (simX[-1][2], np.mean(iters)))

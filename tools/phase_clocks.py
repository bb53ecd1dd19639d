#!/usr/bin/env python
"""Cycles per phase of the one-instance QP kernel (experiment build -DMPCB_PHASE_CLOCKS; never the product library).

    python tools/ab.py build phase=MPCB_PHASE_CLOCKS          # here
    MPCB_LIB_OVERRIDE=mpc_blaster_b200/lib/libmpcb_phase.so python tools/phase_clocks.py [B ...]   # on the GPU box

Prints, per batch size, the clock64() cycles lane 0 of every warp spent in each phase, divided by the number of
(stage, interior-point iteration) pairs, i.e. cycles per stage and iteration as one warp sees them."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_blaster_b200 import BlasterMPC, _lib, scenarios as sc  # noqa: E402

NAMES = ["top of iteration (test, F0 on the first)", "S1 before the factorisation (residuals, P r + p, W product)",
         "S1 factorisation: Gram + Cholesky stages", "S1 factorisation: Householder LQ stages",
         "S1 after the factorisation (substitution, stores)", "S2 forward sweep (affine)", "S3 backward sweep (corrector)",
         "S4 forward sweep (final)", "F4b flat update"]


def main():
    lib = _lib.load()
    out = (C.c_ulonglong * 16)()
    N = 20
    for B in [int(a) for a in sys.argv[1:]] or [1, 1024]:
        mpc = BlasterMPC.canonical(N=N, batch=B, variant=17)
        x0, yref = sc.random_setpoints(B, seed=1234)
        x0 = torch.as_tensor(x0, device="cuda")
        yref = torch.as_tensor(yref, device="cuda")
        trim = torch.as_tensor(sc.hover_trim(6), device="cuda")
        for rep in range(2):
            mpc.reset(x0, trim)
            lib.mpcb_debug_phase_clocks(out, 1)
            mpc.solve(x0, yref, want_traj=False)
            lib.mpcb_debug_phase_clocks(out, 0)
        v = np.array(list(out), dtype=np.float64)
        its, n_gram, n_lq = v[11], v[9], v[10]
        print(f"B = {B}: {its / B:.2f} interior-point iterations per instance (max {int(mpc.iters.max())}); stage factorisations: "
              f"{n_gram / B:.1f} Gram + Cholesky, {n_lq / B:.1f} Householder LQ per instance")
        tot = v[:9].sum()
        for i, name in enumerate(NAMES):
            per = v[i] / (its * N)
            extra = ""
            if i == 2 and n_gram:
                extra = f"   ({v[i] / n_gram:8.0f} per Gram stage)"
            if i == 3 and n_lq:
                extra = f"   ({v[i] / n_lq:8.0f} per LQ stage)"
            print(f"   {name:62s} {per:8.0f} cycles per stage and iteration  {100 * v[i] / tot:5.1f} %{extra}")
        print(f"   {'total':62s} {tot / (its * N):8.0f}      = {tot / B / 1.965e6:.3f} ms per instance at 1.965 GHz")
        del mpc


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Throughput sweep over batch size / horizon / variant (BASELINE.json configs 2, 3, 5), one GPU.
Prints one JSON line per point.  Usage: python tools/sweep.py [--points "B,N,variant;..."]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_blaster_b200 import BlasterMPC, scenarios as sc  # noqa: E402


def run(B, N, variant, scen, steps=5, warmup=2, ws_batch=0, **kw):
    nx, nu = {17: (17, 6), 12: (12, 4), 13: (13, 4)}[variant]
    mpc = BlasterMPC.canonical(N=N, batch=B, variant=variant, ws_batch=ws_batch, **kw)
    if scen == "track":
        x0, yref = sc.lemniscate_tracking(B, N, nx=nx, nu=nu)
    else:
        x0, yref = sc.random_setpoints(B, seed=4567, nx=nx, nu=nu)
    x0 = torch.as_tensor(x0, device="cuda")
    yref = torch.as_tensor(yref, device="cuda")
    trim = torch.as_tensor(sc.hover_trim(nu), device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    mpc.profile(True)
    ts, k2 = [], []
    for i in range(warmup + steps):
        mpc.reset(x0, trim)
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        u0, _, _, st = mpc.solve(x0, yref, want_traj=False)
        b.record()
        torch.cuda.synchronize()
        if i >= warmup:
            ts.append(a.elapsed_time(b))
            k2.append(mpc.last_kernel_ms()[1])
    it = mpc.iters.double()
    out = dict(B=B, N=N, variant=variant, scenario=scen, ms=float(np.mean(ts)), solves_per_s=B / (np.mean(ts) * 1e-3),
               qp_kernel_ms_first_chunk=float(np.mean(k2)), iters_mean=float(it.mean()), iters_max=int(it.max()),
               ok_frac=float((st == 0).double().mean()))
    print(json.dumps(out), flush=True)
    del mpc
    torch.cuda.empty_cache()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", default="1024,20,17,rand;4096,20,17,rand;16384,20,17,rand;65536,20,17,rand;1024,20,12,rand;16384,20,12,rand;"
                                        "65536,40,17,track;16384,80,17,rand")
    ap.add_argument("--qp8-batch", type=int, default=0, help="mpcb_config.qp8_batch (0 = default; huge = never use the four-instances-per-warp kernel)")
    ap.add_argument("--throughput-batch", type=int, default=0, help="mpcb_config.throughput_batch (0 = default)")
    ap.add_argument("--strict", action="store_true", help="mpcb_config.strict_reference")
    a = ap.parse_args()
    for pt in a.points.split(";"):
        B, N, v, scen = pt.split(",")
        run(int(B), int(N), int(v), scen, qp8_batch=a.qp8_batch, throughput_batch=a.throughput_batch, strict_reference=a.strict)

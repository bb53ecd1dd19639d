#!/usr/bin/env python
"""The five BASELINE.json configurations at their full sizes, on 1..8 GPUs of one box.

    python tools/run_configs.py [--configs 1,2,3,4,5] [--scale 1.0] [--out FILE]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_configs.py --configs 4,5

One process per GPU; the batch is sharded contiguously (mpc_blaster_b200.scheduler.shard_range),
there is no collective on the solve path; every number is timed with CUDA events on the launching
stream, max over ranks.  `--scale` shrinks the instance counts (smoke runs).  One JSON line per
measurement on rank 0 (also appended to --out)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_blaster_b200 import BlasterMPC, scenarios as sc  # noqa: E402
from mpc_blaster_b200.scheduler import shard_range  # noqa: E402

RANK = int(os.environ.get("RANK", 0))
WORLD = int(os.environ.get("WORLD_SIZE", 1))
LOCAL = int(os.environ.get("LOCAL_RANK", 0))


def dist_max(v: float) -> float:
    if WORLD == 1:
        return v
    import torch.distributed as dist
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def dist_sum(v: float) -> float:
    if WORLD == 1:
        return v
    import torch.distributed as dist
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])


def barrier():
    torch.cuda.synchronize()
    if WORLD > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


OUT = None


def emit(d):
    d = dict(n_gpus=WORLD, **d)
    if RANK == 0:
        line = json.dumps(d)
        print(line, flush=True)
        if OUT:
            with open(OUT, "a") as f:
                f.write(line + "\n")


def active_set(mpc, X, U, wide_x=False):
    """Bounds at which the new iterate sits (|gap| < 1e-6), per instance: inputs on stages 0..N-1,
    states on stages 1..N-1 (reference blastermodel.py:261-270)."""
    c = mpc.cfg
    nx, nu = mpc.nx, mpc.nu
    lbu = torch.tensor(list(c.lbu)[:nu], device=U.device)
    ubu = torch.tensor(list(c.ubu)[:nu], device=U.device)
    lbx = torch.tensor(list(c.lbx)[:nx], device=U.device)
    ubx = torch.tensor(list(c.ubx)[:nx], device=U.device)
    au = (((U - lbu).abs() < 1e-6) | ((U - ubu).abs() < 1e-6)).sum(dim=(1, 2)).double()
    Xi = X[:, 1:-1]
    ax = (((Xi - lbx).abs() < 1e-6) | ((Xi - ubx).abs() < 1e-6)).sum(dim=(1, 2)).double()
    return au, ax


def timed_solves(mpc, x0, yref, trim, steps, warmup, want_traj=False, keep=None):
    """steps timed solves of the local shard from the same cold iterate; returns (mean ms, last outputs)."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    # inputs resident in HBM before the timed region (config 3's per-stage yref is 0.5 GB: converting it inside the timed events
    # added a pageable host-to-device copy of 40 ... 150 ms to every step of the round-1 and first round-2 tables)
    x0, yref, trim = (torch.as_tensor(a, dtype=torch.float64, device="cuda") for a in (x0, yref, trim))
    ts, out = [], None
    for i in range(warmup + steps):
        mpc.reset(x0, trim)
        flush.zero_()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = mpc.solve(x0, yref, want_traj=want_traj)
        b.record()
        torch.cuda.synchronize()
        if i >= warmup:
            ts.append(dist_max(a.elapsed_time(b)))
    return float(np.mean(ts)), out


def solver_stats(mpc, st):
    it = mpc.iters.double()
    n = float(st.numel())
    tot = dist_sum(n)
    return dict(mean_ipm_iters=dist_sum(float(it.sum())) / tot, max_ipm_iters=int(dist_max(float(it.max()))),
                converged_frac=dist_sum(float((st == 0).sum())) / tot)


def config1():
    """Single BLASTER quadrotor, hover to set-point (simulation_blaster.py:47-48), N=20, closed loop."""
    if RANK != 0:
        return
    x0, yref = sc.hover_to_setpoint()
    mpc = BlasterMPC.canonical(N=20, batch=1)
    x = torch.as_tensor(x0, device="cuda")
    yref = yref[0]
    yr = torch.as_tensor(yref, device="cuda")
    ts, its = [], []
    for s in range(200):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        u0, _, _, st = mpc.solve(x, yr, want_traj=False)
        b.record()
        torch.cuda.synchronize()
        assert int(st[0]) == 0
        ts.append(a.elapsed_time(b))
        its.append(int(mpc.iters[0]))
        x = mpc.step_plant(x, u0)
    d = dict(config="1: single instance, hover-to-setpoint closed loop, N=20, 200 control steps", gpu_ms_first_solve=ts[0],
             gpu_ms_p50=float(np.percentile(ts[1:], 50)), gpu_ms_p99=float(np.percentile(ts[1:], 99)), mean_ipm_iters=float(np.mean(its)),
             final_z=float(x[0, 2]), realtime_budget_ms=1000.0 / 30)
    try:
        from oracle import blaster_oracle as bo, c_oracle as co
        c = co.BatchRTI(bo.canonical_problem(20), 1, nthreads=1)
        xc = x0.copy()
        tc = []
        for s in range(60):
            t = time.perf_counter()
            u, _, _, _ = c.solve(xc, yref)
            tc.append(1e3 * (time.perf_counter() - t))
            xc = co.plant_step(c.P, xc, u)
        d.update(cpu_port_ms_first_solve=tc[0], cpu_port_ms_p50=float(np.percentile(tc[1:], 50)), cpu_threads=1)
    except Exception as e:  # the oracle is optional here
        d["cpu_port"] = f"unavailable: {e}"
    emit(d)
    del mpc


def config2(scale):
    B = max(WORLD, int(1024 * scale))
    for variant in (17, 12):
        nx, nu = (17, 6) if variant == 17 else (12, 4)
        lo, hi = shard_range(B, RANK, WORLD)
        x0, yref = sc.random_setpoints(B, seed=1234, nx=nx, nu=nu)
        mpc = BlasterMPC.canonical(N=20, batch=hi - lo, variant=variant)
        ms, (u0, _, _, st) = timed_solves(mpc, x0[lo:hi], yref[lo:hi], sc.hover_trim(nu), 20, 5)
        emit(dict(config="2: batch of 1,024 instances, randomised x0/x_ref, N=20", variant=variant, batch=B, horizon=20, ms=ms,
                  solves_per_s=B / ms * 1e3, **solver_stats(mpc, st)))
        del mpc


def config3(scale):
    B, N = max(WORLD, int(65536 * scale)), 40
    lo, hi = shard_range(B, RANK, WORLD)
    x0, yref = sc.lemniscate_tracking(B, N)
    x0, yref = x0[lo:hi], yref[lo:hi]
    for name, widen in (("3a: reference bounds (input and state bounds bind)", 1.0), ("3b: state bounds widened x1e3 (only input bounds can bind)", 1e3)):
        kw = {}
        if widen != 1.0:
            ref = BlasterMPC.canonical(N=2, batch=1).cfg
            kw["statesBound"] = np.array([list(ref.lbx), list(ref.ubx)]) * widen
        mpc = BlasterMPC.canonical(N=N, batch=hi - lo, **kw)
        ms, (u0, X, U, st) = timed_solves(mpc, x0, yref, sc.hover_trim(), 3, 1, want_traj=True)
        ok = st == 0
        au, ax = active_set(mpc, X, U)
        nok = dist_sum(float(ok.sum()))
        emit(dict(config="3: batch of 65,536 instances, aggressive figure-eight tracking, per-stage yref, N=40 -- " + name, batch=B, horizon=N,
                  ms=ms, solves_per_s=B / ms * 1e3, mean_active_input_bounds=dist_sum(float(au[ok].sum())) / max(nok, 1),
                  mean_active_state_bounds=dist_sum(float(ax[ok].sum())) / max(nok, 1), input_bounds_per_instance=N * 6,
                  state_bounds_per_instance=(N - 1) * 17, **solver_stats(mpc, st)))
        del mpc, X, U
        torch.cuda.empty_cache()


def config4(scale, steps=500):
    B, N = max(WORLD, int(16384 * scale)), 20
    lo, hi = shard_range(B, RANK, WORLD)
    x0, yref = sc.closed_loop_setpoints(B, seed=3456)
    x0, yref = x0[lo:hi], yref[lo:hi]
    mpc = BlasterMPC.canonical(N=N, batch=hi - lo)
    mpc.reset(x0, sc.hover_trim())
    mpc.closed_loop(x0, yref, steps=2)  # warm-up (module load, allocator)
    mpc.reset(x0, sc.hover_trim())
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    xT, u_last, n_fail, iters = mpc.closed_loop(x0, yref, steps=steps)
    b.record()
    torch.cuda.synchronize()
    ms = dist_max(a.elapsed_time(b))
    err = (xT[:, :3] - torch.as_tensor(yref[:, :3], device="cuda")).norm(dim=1)
    emit(dict(config="4: closed-loop Monte-Carlo, 16,384 quadrotors x 500 control steps, un-shifted warm-started SQP-RTI, plant on device",
              batch=B, steps=steps, horizon=N, seconds=ms * 1e-3, instance_steps_per_s=B * steps / ms * 1e3,
              failed_solves=int(dist_sum(float(n_fail.sum()))), mean_ipm_iters=dist_sum(float(iters.double().sum())) / (B * steps),
              mean_final_position_error_m=dist_sum(float(err.sum())) / B))
    del mpc


def config5(scale, horizons=(20, 40, 80), variants=(17, 12)):
    B = max(WORLD, int(1048576 * scale))
    lo, hi = shard_range(B, RANK, WORLD)
    for variant in variants:
        nx, nu = (17, 6) if variant == 17 else (12, 4)
        x0, yref = sc.random_setpoints(B, seed=4567, nx=nx, nu=nu)
        x0, yref = x0[lo:hi], yref[lo:hi]
        for N in horizons:
            mpc = BlasterMPC.canonical(N=N, batch=hi - lo, variant=variant)
            ms, (u0, _, _, st) = timed_solves(mpc, x0, yref, sc.hover_trim(nu), 2, 1)
            emit(dict(config="5: throughput sweep, 1M instances (FP64; FP32 is not implemented, DESIGN.md)", variant=variant, batch=B, horizon=N,
                      ms=ms, solves_per_s=B / ms * 1e3, **solver_stats(mpc, st)))
            del mpc
            torch.cuda.empty_cache()


def main():
    global OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps4", type=int, default=500)
    ap.add_argument("--horizons5", default="20,40,80")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    OUT = a.out
    torch.cuda.set_device(LOCAL)
    if WORLD > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
    want = [int(c) for c in a.configs.split(",")]
    if 1 in want:
        config1()
    if 2 in want:
        config2(a.scale)
    if 3 in want:
        config3(a.scale)
    if 4 in want:
        config4(a.scale, a.steps4)
    if 5 in want:
        config5(a.scale, tuple(int(h) for h in a.horizons5.split(",")))
    if WORLD > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

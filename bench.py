#!/usr/bin/env python
"""Headline benchmark: batched BLASTER MPC solves/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): 1,024 independent BLASTER17 instances per GPU,
horizon N=20, randomised x0 / set-point (mpc_blaster_b200.scenarios.random_setpoints), SQP
iterate initialised at (x0, hover trim), cold-started QP.  One "step" = one SQP-RTI solve of
the whole batch (rollout+sensitivities kernel, then the Riccati-IPM kernel).  The iterate is
re-initialised and L2 is flushed between steps, outside the timed events, so every timed
step does the same work from a cold cache.  Multi-GPU: one process per GPU, each with its
own 1,024-instance batch (weak scaling), no collective on the solve path; one final gather.

`value` is timed with CUDA events on the launching stream with inputs resident in HBM;
`e2e` goes through the C ABI's host entry point (mpcb_solve_host: pinned staging, H2D,
solve, D2H) with NumPy buffers.  `cpu_baseline` / `--impl reference` time the C oracle
(oracle/mpc_oracle.c, kind "port": acados/HPIPM cannot be built here) on the host cores.

Beside the headline the line carries: `roofline` (the binding resource of the dominant kernel: the
FP64 FMA pipe, achieved algorithmic TFLOP/s against the DFMA peak measured in the same run) and
`roofline_hbm` (algorithmic bytes against MEASURED_PEAKS.json, with the ncu DRAM traffic);
`solver.p50_ms / p99_ms` over >= 200 launches whatever --steps is; the explicit KKT residuals of
the last solve's solutions; `zero_iterate` (acados' default all-zero initial iterate instead of
(x0, hover trim)); `large_batch` (65,536 instances per GPU of BLASTER17 and QUAD12 on the
four-instances-per-warp kernel, the kernel configs 3-5 run on, with its own roofline).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mpc_blaster_b200 import scenarios as sc  # noqa: E402

HORIZON = 20
BATCH = 1024
VARIANT = 17
METRIC = "MPC solves/sec (N=20, batched)"
UNIT = "solves/s"


def workload(rank: int, batch: int = BATCH, variant: int = VARIANT):
    nx, nu = (17, 6) if variant == 17 else (12, 4)
    x0, yref = sc.random_setpoints(batch, seed=1234 + rank, nx=nx, nu=nu)
    return x0, yref, sc.hover_trim(nu)


def config_dict(world: int, batch: int, horizon: int, variant: int):
    return {"workload": f"configs[1]: batch of {batch} independent quadrotor MPC instances per GPU (randomised x0/x_ref), "
                        f"N={horizon}, BLASTER{variant} ({'17 states / 6 inputs' if variant == 17 else '12 states / 4 inputs'}), "
                        "one SQP-RTI iteration, HPIPM-default KKT tolerances",
            "batch_per_gpu": batch, "horizon": horizon, "variant": variant, "global_batch": batch * world,
            "parallelism": f"dp{world} (independent instances, no collective on the solve path)",
            "l2_flush_between_steps": True, "iterate_reset_between_steps": True}


# ---------------------------------------------------------------------------- flops / bytes model
def algorithmic_bytes_per_solve(nx, nu, N, s=8):
    """read x0, yref, p; read + write the iterate."""
    return s * ((nx + (nx + nu) + 25) + 2 * ((N + 1) * nx + N * nu))


def algorithmic_flops_per_solve(nx, nu, N, n_fact):
    nnzA, c_fJ = (59, 375) if nx == 17 else (36, 230)
    nz = nx + nu
    f_lin = 4 * (2 * nnzA * nz + c_fJ) + 8 * nx * (nz + 1)
    f_ric = nx * nx * nz + nx * nz * nz + nz ** 3 / 3.0
    f_sub = 2 * nz * nz + 4 * nx * nz
    return N * f_lin + n_fact * N * f_ric + 2 * n_fact * N * f_sub


# ---------------------------------------------------------------------------- clocks sampler
class ClockSampler(threading.Thread):
    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_steps(steps: int, warmup: int, batch: int, horizon: int, variant: int):
    """Time the C oracle on all host cores: each step = the same 1,024-instance batch."""
    from oracle import blaster_oracle as bo
    from oracle import c_oracle as co
    P = bo.canonical_problem(horizon, variant)
    x0, yref, trim = workload(0, batch, variant)
    # all host cores, whatever OMP_NUM_THREADS says (torchrun sets it to 1)
    orc = co.BatchRTI(P, batch, nthreads=len(os.sched_getaffinity(0)))
    times = []
    for i in range(warmup + steps):
        orc.reset(x0, trim)
        t = time.perf_counter()
        orc.solve(x0, yref)
        dt = time.perf_counter() - t
        if i >= warmup:
            times.append(dt)
    ok = float((orc.status == 0).mean())
    return times, orc.nthreads, float(orc.iters.mean()), ok


def run_reference(args):
    """`--impl reference`: the reference's CPU path for this solve.  acados/HPIPM/CasADi are not
    installable here, so this is the C oracle port with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    times, cores, iters, ok = cpu_steps(args.steps, args.warmup, BATCH, HORIZON, VARIANT)
    total = sum(times)
    value = BATCH * len(times) / total
    sample = f"{BATCH} instances (the full per-GPU batch) per step, {len(times)} steps"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(world, BATCH, HORIZON, VARIANT),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "mean_ipm_iters": iters, "converged_frac": ok},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


# ---------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from mpc_blaster_b200 import BlasterMPC

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, N = args.batch, HORIZON
    mpc = BlasterMPC.canonical(N=N, batch=B, variant=VARIANT)
    nx, nu = mpc.nx, mpc.nu
    x0_h, yref_h, trim_h = workload(rank, B, VARIANT)
    x0 = torch.as_tensor(x0_h, device=dev)
    yref = torch.as_tensor(yref_h, device=dev)
    trim = torch.as_tensor(trim_h, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # 256 MiB > 126 MB L2
    mpc.profile(True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    for _ in range(args.warmup):
        mpc.reset(x0, trim)
        flush.zero_()
        mpc.solve(x0, yref, want_traj=False)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    k1_ms, k2_ms = [], []
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = mpc.kernel_launches()
    launches = 0
    for a, b in ev:
        mpc.reset(x0, trim)
        flush.zero_()
        la = mpc.kernel_launches()
        a.record()
        u0, _, _, status = mpc.solve(x0, yref, want_traj=False)
        b.record()
        launches += mpc.kernel_launches() - la
        t1, t2 = mpc.last_kernel_ms()
        k1_ms.append(t1)
        k2_ms.append(t2)
    barrier()
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    iters_mean = float(mpc.iters.double().mean().item())
    ok_frac = float((status == 0).double().mean().item())

    # ---- end to end through the host entry point of the C ABI
    for _ in range(max(1, args.warmup // 2)):
        mpc.reset(x0, trim)
        mpc.solve_host(x0_h, yref_h)
    barrier()
    e2e_t = 0.0
    for _ in range(args.steps):
        mpc.reset(x0, trim)
        flush.zero_()
        torch.cuda.synchronize()
        t = time.perf_counter()
        u0_h, _, _, st_h = mpc.solve_host(x0_h, yref_h)
        e2e_t += time.perf_counter() - t
    e2e_all = torch.tensor([e2e_t], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_all, op=dist.ReduceOp.MAX)
    e2e_t = float(e2e_all.item())
    assert np.array_equal(u0_h, u0.cpu().numpy()), "host path and device path disagree"

    # ---- the QUAD12 instantiation (12 states / 4 inputs, north_star's sizing) on the same scenario, N=1 only
    quad12 = None
    if world == 1 and not args.no_quad12:
        q = BlasterMPC.canonical(N=N, batch=B, variant=12)
        qx_h, qy_h, qt_h = workload(rank, B, 12)
        qx, qy, qt = (torch.as_tensor(a, device=dev) for a in (qx_h, qy_h, qt_h))
        qms = []
        for i in range(args.warmup + max(5, args.steps // 3)):
            q.reset(qx, qt)
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _, _, _, qst = q.solve(qx, qy, want_traj=False)
            b.record()
            torch.cuda.synchronize()
            if i >= args.warmup:
                qms.append(a.elapsed_time(b))
        quad12 = {"value": B / (float(np.mean(qms)) * 1e-3), "unit": UNIT, "ms_per_step": float(np.mean(qms)),
                  "mean_ipm_iters": float(q.iters.double().mean().item()), "converged_frac": float((qst == 0).double().mean().item())}
        del q

    # ---- per-launch latency over >= 200 launches, whatever --steps is
    lat_n = max(200, args.steps)
    lev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(lat_n)]
    for a, b in lev:
        mpc.reset(x0, trim)
        flush.zero_()
        a.record()
        mpc.solve(x0, yref, want_traj=False)
        b.record()
    torch.cuda.synchronize()
    lat_ms = np.array([a.elapsed_time(b) for a, b in lev])

    # ---- explicit KKT residuals of the solutions of the last solve (evaluated from the exported QP data)
    from mpc_blaster_b200 import diagnostics
    kkt = {k: float(v[status == 0].max().item()) for k, v in diagnostics.explicit_kkt_residuals(mpc, B).items()}
    max_iters_rank = torch.tensor([int(mpc.iters.max().item())], dtype=torch.int64, device=dev)
    if world > 1:
        gathered = [torch.zeros_like(max_iters_rank) for _ in range(world)]
        dist.all_gather(gathered, max_iters_rank)
        max_iters_rank = torch.cat(gathered)
    max_iters_rank = [int(v) for v in max_iters_rank.tolist()]

    # ---- acados' default initial iterate (all zero) instead of (x0, hover trim)
    zms = []
    for i in range(3 + 10):
        mpc.reset()
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _, _, _, zst = mpc.solve(x0, yref, want_traj=False)
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            zms.append(a.elapsed_time(b))
    zero_iterate = {"value": B / (float(np.mean(zms)) * 1e-3), "unit": UNIT + " per GPU", "ms_per_step": float(np.mean(zms)),
                    "mean_ipm_iters": float(mpc.iters.double().mean().item()), "max_ipm_iters": int(mpc.iters.max().item()),
                    "converged_frac": float((zst == 0).double().mean().item()),
                    "note": "first solve from the all-zero iterate (acados' default when nothing is set): the linearisation point is "
                            "the origin, metres away from x0"}
    fp64_peak_early = mpc.fp64_peak_tflops()
    del mpc

    # ---- the large-batch kernel (qp8_kernel: four instances per warp), the one configs 3-5 run on
    large = None
    if not args.no_large:
        large = {}
        LB = args.large_batch
        for variant in (17, 12):
            lm = BlasterMPC.canonical(N=N, batch=LB, variant=variant)
            lx_h, ly_h, lt_h = workload(rank, LB, variant)
            lx, ly, lt = (torch.as_tensor(a, device=dev) for a in (lx_h, ly_h, lt_h))
            lm.profile(True)
            lms, lk2 = [], []
            for i in range(2 + 5):
                lm.reset(lx, lt)
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                lu0, _, _, lst = lm.solve(lx, ly, want_traj=False)
                b.record()
                torch.cuda.synchronize()
                if i >= 2:
                    lms.append(a.elapsed_time(b))
                    lk2.append(lm.last_kernel_ms()[1])
            t = torch.tensor([sum(lms)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            lit = float(lm.iters.double().mean().item())
            lnx, lnu = lm.nx, lm.nu
            lflops = algorithmic_flops_per_solve(lnx, lnu, N, lit) - algorithmic_flops_per_solve(lnx, lnu, N, 0)
            k2s = float(np.mean(lk2)) * 1e-3
            ent = {"value": LB * world * len(lms) / (float(t.item()) * 1e-3), "unit": UNIT, "batch_per_gpu": LB, "ms_per_step": float(t.item()) / len(lms),
                   "kernel": f"qp8_kernel<{lnx},{lnu}>", "kernel_ms": k2s * 1e3, "mean_ipm_iters": lit,
                   "converged_frac": float((lst == 0).double().mean().item()),
                   "roofline": {"bound": "fp64_fma", "achieved": lflops * LB / k2s / 1e12, "peak": fp64_peak_early, "unit": "TFLOP/s",
                                "frac": lflops * LB / k2s / 1e12 / fp64_peak_early},
                   "roofline_hbm": {"achieved": algorithmic_bytes_per_solve(lnx, lnu, N) * LB / k2s / 1e9, "unit": "GB/s",
                                    "algorithmic_bytes_per_launch": algorithmic_bytes_per_solve(lnx, lnu, N) * LB}}
            if world == 1 and not args.no_cpu:
                # spot check of this very launch against the C oracle (64 sampled instances)
                from oracle import blaster_oracle as bo
                from oracle import c_oracle as co
                idx = np.sort(np.random.default_rng(7).choice(LB, 64, replace=False))
                orc = co.BatchRTI(bo.canonical_problem(N, variant), len(idx), nthreads=len(os.sched_getaffinity(0)))
                orc.reset(lx_h[idx], lt_h)
                uo, _, _, sto = orc.solve(lx_h[idx], ly_h[idx])
                ti = torch.as_tensor(idx, device=dev)
                okk = sto == 0
                ent["oracle_spot_check"] = {"instances": len(idx), "status_equal": bool((lst[ti].cpu().numpy() == sto).all()),
                                            "iters_equal": bool((lm.iters[ti].cpu().numpy() == orc.iters).all()),
                                            "max_du0": float(np.abs(lu0[ti].cpu().numpy()[okk] - uo[okk]).max())}
            large[f"blaster{variant}" if variant == 17 else f"quad{variant}"] = ent
            del lm
            torch.cuda.empty_cache()

    # ---- final result gather (the only collective of the job)
    if world > 1:
        from mpc_blaster_b200.scheduler import gather_batch
        all_u0 = gather_batch(u0, B * world)
        assert all_u0.shape == (B * world, nu)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = B * world * args.steps / (total_ms * 1e-3)
    e2e_value = B * world * args.steps / e2e_t
    k2 = float(np.mean(k2_ms)) * 1e-3
    k1 = float(np.mean(k1_ms)) * 1e-3
    # roofline of the dominant kernel (qp_kernel)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback B200_PROFILING.md"
    alg_bytes = algorithmic_bytes_per_solve(nx, nu, N) * B
    achieved_gbs = alg_bytes / k2 / 1e9
    fp64_peak = fp64_peak_early
    flops = algorithmic_flops_per_solve(nx, nu, N, iters_mean)
    f_lin_only = algorithmic_flops_per_solve(nx, nu, N, 0)
    qp_tflops = (flops - f_lin_only) * B / k2 / 1e12
    lin_tflops = f_lin_only * B / k1 / 1e12
    traffic, ncu = None, {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "qp_kernel_traffic.json")))
        traffic = ncu.get("dram_bytes_per_launch")
    except Exception:
        pass

    # ---- CPU baseline on a bounded sample of the same workload (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        reps = max(2, args.cpu_steps)
        times, cores, cpu_iters, cpu_ok = cpu_steps(reps, 1, B, N, VARIANT)
        cpu = {"value": B * len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"the same {B}-instance batch, {len(times)} repeats ({sum(times):.1f} s of wall time on {cores} threads)",
               "mean_ipm_iters": cpu_iters, "converged_frac": cpu_ok}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(world, B, N, VARIANT),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(x0_h.nbytes + yref_h.nbytes),
                    "d2h_bytes_per_step": int(B * nu * 8 + 2 * B * 4)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64_fma", "kernel": "qp_kernel<17,6>", "achieved": qp_tflops, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": qp_tflops / fp64_peak, "traffic": traffic,
                         "peak_source": "DFMA micro-kernel in this run (mpcb_fp64_peak); MEASURED_PEAKS.json carries no FP64 figure",
                         "algorithmic_flops_per_launch": (flops - f_lin_only) * B, "kernel_ms": k2 * 1e3,
                         "ncu_pipe_fp64_cycles_active_pct": ncu.get("sm__pipe_fp64_cycles_active_pct"),
                         "ncu_issue_active_pct": ncu.get("smsp__issue_active_pct"),
                         "note": "SURVEY 8(d): the path is bound by the FP64 CUDA-core pipe and the dependent chain of the stage "
                                 "factorisation, not by HBM or the tensor cores (no FP64 tcgen05 kind)"},
            "roofline_hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                             "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes},
            "fp64": {"qp_kernel_tflops": qp_tflops, "linearize_kernel_tflops": lin_tflops, "peak_tflops_measured": fp64_peak,
                     "qp_frac": qp_tflops / fp64_peak, "linearize_kernel_ms": k1 * 1e3,
                     "algorithmic_mflop_per_solve": flops / 1e6},
            "solver": {"mean_ipm_iters": iters_mean, "converged_frac": ok_frac, "max_ipm_iters_per_rank": max_iters_rank,
                       "p50_ms": float(np.percentile(lat_ms, 50)), "p99_ms": float(np.percentile(lat_ms, 99)), "latency_launches": int(lat_n),
                       "max_explicit_res_stat": kkt["stat"], "max_explicit_res_eq": kkt["eq"],
                       "max_explicit_res_ineq": kkt["ineq"], "max_explicit_res_comp": kkt["comp"]},
            "zero_iterate": zero_iterate, "large_batch": large,
            "cpu_baseline": cpu, "quad12": quad12}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _emit(line: dict):
    """The ONE JSON line of the contract, on the process's original stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def main():
    # Native libraries print to fd 1 (NCCL's version banner under torchrun, for one): keep the
    # original stdout for the JSON line alone and send everything else to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-quad12", action="store_true")
    ap.add_argument("--no-large", action="store_true", help="skip the 65,536-instance leg on the four-instances-per-warp kernel")
    ap.add_argument("--large-batch", type=int, default=65536)
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "graft" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

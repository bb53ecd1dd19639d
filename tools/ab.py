#!/usr/bin/env python
"""A/B timing of experimental builds of the CUDA library (compile-time -D variants).

    python tools/ab.py build  name1=DEF1,DEF2 name2= ...      # here (nvcc, no GPU): lib/libmpcb_<name>.so
    python tools/ab.py run    name1 name2 ... [--points "B,N,variant,scen;..."]   # on the GPU box

`run` starts one sweep process per (build, point) with MPCB_LIB_OVERRIDE set, so every
variant is timed in the same gpurun call on the same GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    mode, args = sys.argv[1], sys.argv[2:]
    from mpc_blaster_b200 import _build
    if mode == "build":
        for a in args:
            name, _, defs = a.partition("=")
            out = os.path.join(_build.LIB_DIR, f"libmpcb_{name}.so")
            _build.build(defines=[d for d in defs.split(",") if d], out=out)
            print("built", out)
        return
    points = "1024,20,17,rand;16384,20,17,rand"
    if "--points" in args:
        i = args.index("--points")
        points = args[i + 1]
        args = args[:i] + args[i + 2:]
    for pt in points.split(";"):
        for name in args:
            env = dict(os.environ, MPCB_LIB_OVERRIDE=os.path.join(_build.LIB_DIR, f"libmpcb_{name}.so"))
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sweep.py"), "--points", pt], env=env, capture_output=True, text=True)
            try:
                d = json.loads(r.stdout.strip().splitlines()[-1])
                print(f"{name:>12s}  B={d['B']:<7d} N={d['N']:<3d} v={d['variant']:<3d} {d['scenario']:5s} {d['ms']:9.3f} ms  {d['solves_per_s']:10.0f} solves/s  "
                      f"qp {d['qp_kernel_ms_first_chunk']:8.3f} ms  ok {d['ok_frac']:.4f}", flush=True)
            except Exception:
                print(name, pt, "FAILED", r.stdout[-300:], r.stderr[-600:], flush=True)


if __name__ == "__main__":
    main()

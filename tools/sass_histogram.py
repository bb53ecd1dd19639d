#!/usr/bin/env python
"""SASS opcode histogram of libmpcb.so (cuobjdump -sass), whole library and per kernel, with the mnemonics that prove the
Blackwell-specific paths: DMMA (FP64 tensor-core MMA), UBLKCP / UBLKPF (TMA bulk copy / L2 prefetch), SYNCS (mbarrier).

    python tools/sass_histogram.py [lib] > profiles/r02_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mpc_blaster_b200", "lib", "libmpcb.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = re.search(r"arch = (\S+)", out)
per, cur, total = collections.OrderedDict(), None, collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0].replace("void ", "")
        cur = per.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", line)
    if m and cur is not None:
        op = m.group(1).split(".")[0]
        cur[op] += 1
        total[op] += 1
print(f"# SASS opcode histogram of `{os.path.relpath(lib, ROOT)}` ({arch.group(1) if arch else '?'}; `cuobjdump -sass`, `tools/sass_histogram.py`)\n")
print("Whole library, top 30:\n\n| opcode | count |\n|---|---:|")
for op, n in total.most_common(30):
    print(f"| {op} | {n} |")
keys = ["DFMA", "DMUL", "DADD", "DMMA", "MUFU", "LDS", "STS", "LDG", "STG", "SHFL", "UBLKCP", "UBLKPF", "SYNCS", "BAR", "HMMA", "UTCMMA", "LDTM", "UTMALDG"]
print("\nPer kernel (instructions in the binary, not executed counts):\n\n| kernel | total | " + " | ".join(keys) + " |\n|---|---:|" + "---:|" * len(keys))
for name, c in per.items():
    if sum(c.values()) < 200:
        continue
    print(f"| `{name}` | {sum(c.values())} | " + " | ".join(str(c.get(k, 0)) for k in keys) + " |")
print("\nDMMA = `mma.sync.aligned.m8n8k4.f64` (FP64 tensor cores: the `W = [B A]'L` product and the Gram matrix of the latency "
      "variant of `qp_kernel`); UBLKCP / UBLKPF = `cp.async.bulk` (TMA bulk copy into shared memory / L2 prefetch) and SYNCS = "
      "mbarrier operations of the stage-prefetch pipelines; no HMMA / UTCMMA / LDTM / UTMALDG: the path is FP64 and FP64 has no "
      "`tcgen05` kind.")

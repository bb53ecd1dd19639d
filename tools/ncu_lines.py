#!/usr/bin/env python
"""Join an ncu report's per-SASS-instruction samples with nvdisasm line info and print a
per-source-line profile of one kernel (samples, stall mix, executed instructions).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep mpc_blaster_b200/lib/libmpcb.so qp_kernelILi17ELi6ELi1ELi2ELi1E

The pattern must select ONE function of the library (mangled-name substring; several template instantiations
share a prefix): with more than one match the tool stops and lists them -- their line tables would overwrite
each other.  Run it from a checkout of the sources the library was built from.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    rep, so, pat = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # a report may hold several kernels: take the section whose "Kernel Name" row matches
    want = pat.split("ILi")[0]
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    sec = next((i for i in starts if want in rows[i][1]), starts[0] if starts else 0)
    hi = next(i for i in range(sec, len(rows)) if rows[i] and rows[i][0] == "Address")
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    inst = []
    for r in rows[hi + 1:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) == len(hdr):
            inst.append(r)
    base = int(inst[0][0], 16)
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    line_of = {}
    cur, infn = None, False
    matches = [l.strip() for l in sass.splitlines() if l.startswith(".text.") and pat in l]
    if len(matches) != 1:
        sys.exit(f"pattern {pat!r} selects {len(matches)} functions, need exactly one:\n  " + "\n  ".join(matches))
    for l in sass.splitlines():
        if l.startswith(".text."):
            infn = pat in l
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
        if m:
            line_of[int(m.group(1), 16)] = cur
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = defaultdict(lambda: defaultdict(float))
    tot = 0
    for r in inst:
        off = int(r[0], 16) - base
        key = line_of.get(off, ("?", 0))
        s = float(r[col["# Samples"]] or 0)
        agg[key]["samples"] += s
        agg[key]["inst"] += float(r[col["Instructions Executed"]] or 0)
        for h in stalls:
            agg[key][h] += float(r[col[h]] or 0)
        tot += s
    print(f"total samples {tot:.0f}, instructions {sum(a['inst'] for a in agg.values()):.3g}")
    src_cache = {}
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        f, ln = key
        text = ""
        for root in ("mpc_blaster_b200/csrc", "."):
            p = os.path.join(root, f)
            if os.path.exists(p):
                src_cache.setdefault(p, open(p).read().splitlines())
                if 0 < ln <= len(src_cache[p]):
                    text = src_cache[p][ln - 1].strip()[:70]
                break
        mix = sorted(((a[h], h[6:]) for h in stalls), reverse=True)[:3]
        mixs = " ".join(f"{n}:{100 * v / max(a['samples'], 1):.0f}%" for v, n in mix)
        print(f"{100 * a['samples'] / tot:5.1f}%  inst {a['inst'] / 1024:8.0f}/warp  {f}:{ln:<4d} [{mixs}]  {text}")


if __name__ == "__main__":
    main()

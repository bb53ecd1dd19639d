#!/usr/bin/env python
"""Turn the raw ncu outputs of a round (gpurun_out/) into the tracked summaries under profiles/.

    python tools/summarize_profiles.py <tag> <launches.csv> <full.ncu-rep> ["workload text"] [--bench]

--bench: the capture is of bench.py's own workload (configs[1]); only then is profiles/qp_kernel_traffic.json,
the per-launch DRAM traffic bench.py reports as roofline.traffic, rewritten.
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__shared_mem_per_block_static", "smsp__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def raw_only(name, csv_path, workload):
    """`--raw-only <name> <raw.csv> "workload"`: summary of a capture whose report stayed on the GPU box."""
    rr = list(csv.reader(open(csv_path)))
    hdr = rr[0]
    with open(os.path.join(ROOT, "profiles", f"{name}_ncu_summary.md"), "w") as f:
        f.write(f"# {name}: `ncu --set full --clock-control none` ({workload})\n\n")
        for r in rr[2:]:
            kn = r[hdr.index("Kernel Name")].split("(")[0]
            f.write(f"## `{kn}`\n\n| metric | unit | value |\n|---|---|---:|\n")
            vals = {}
            for i, h in enumerate(hdr):
                if h in KEYS or ("pcsamp_warps_issue_stalled" in h and "not_issued" not in h):
                    f.write(f"| {h} | {rr[1][i]} | {r[i]} |\n")
                    vals[h] = (r[i], rr[1][i])
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
            tb = sum(float(vals[k][0]) * mult[vals[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            f.write(f"\nDRAM traffic per launch: {tb / 1e6:.1f} MB\n\n")
            print(name, kn, f"{float(vals['gpu__time_duration.sum'][0]):.3f} {vals['gpu__time_duration.sum'][1]}", f"DRAM {tb / 1e9:.2f} GB",
                  "L2 hit", vals.get("lts__t_sector_hit_rate.pct", ("?",))[0], "issue", vals.get("smsp__issue_active.avg.pct_of_peak_sustained_active", ("?",))[0])


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--raw-only":
        return raw_only(*sys.argv[2:5])
    args = [a for a in sys.argv[1:] if a != "--bench"]
    is_bench = "--bench" in sys.argv[1:]
    tag, launches, rep = args[:3]
    workload = args[3] if len(args) > 3 else "B=1,024, N=20, BLASTER17"
    out_dir = os.path.join(ROOT, "profiles")
    os.makedirs(out_dir, exist_ok=True)
    # ---- launch list
    rows = list(csv.reader(open(launches)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    col = {h: i for i, h in enumerate(rows[hi])}
    agg = defaultdict(lambda: [0, 0.0])
    seq = []
    for r in rows[hi + 1:]:
        if len(r) != len(rows[hi]):
            continue
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        ns = float(r[col["Metric Value"]])
        agg[name][0] += 1
        agg[name][1] += ns
        seq.append((name, ns))
    def ours(k):
        return "at::" not in k and "fp64_peak" not in k
    total_ours = sum(v[1] for k, v in agg.items() if ours(k))
    with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: kernel launch list of `bench.py` (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | mean us | share of this library's kernel time |\n|---|---:|---:|---:|\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            share = f"{100 * ns / total_ours:.1f} %" if ours(k) else ("(FP64 peak micro-benchmark, outside the timed steps)" if "fp64_peak" in k else "(torch: L2 flush / bookkeeping, outside the timed events)")
            f.write(f"| `{k}` | {n} | {ns / n / 1e3:.1f} | {share} |\n")
        f.write("\nFirst launches in order:\n\n```\n")
        for name, ns in seq[:16]:
            f.write(f"{ns / 1e3:10.1f} us  {name}\n")
        f.write("```\n")
    # ---- full capture
    # `rep` is an .ncu-rep, or the CSV of its raw page exported on the GPU box (`ncu -i rep --page raw --csv`): the reports of
    # the large-batch captures are too big to bring back whole
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    hdr = rr[0]
    traffic, pipes = {}, {}
    with open(os.path.join(out_dir, f"{tag}_ncu_summary.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none` of the two hot kernels ({workload})\n\n")
        for r in rr[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0]
            f.write(f"## `{name}`\n\n| metric | unit | value |\n|---|---|---:|\n")
            vals = {}
            for i, h in enumerate(hdr):
                if h in KEYS or ("pcsamp_warps_issue_stalled" in h and "not_issued" not in h):
                    f.write(f"| {h} | {rr[1][i]} | {r[i]} |\n")
                    vals[h] = (r[i], rr[1][i])

            def gb(k):
                v, u = vals[k]
                return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
            tb = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
            traffic[name] = tb
            pipes[name] = {k: float(vals[k][0]) for k in ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                                                          "smsp__issue_active.avg.pct_of_peak_sustained_active",
                                                          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                                                          "sm__warps_active.avg.pct_of_peak_sustained_active",
                                                          "lts__t_sector_hit_rate.pct") if k in vals}
            f.write(f"\nDRAM traffic per launch: {tb / 1e6:.1f} MB\n\n")
    if is_bench:
        qk = next(k for k in traffic if "qp_kernel" in k)
        pp = pipes[qk]
        json.dump({"kernel": "qp_kernel<17,6,1,2,1,false>", "dram_bytes_per_launch": traffic[qk], "source": f"profiles/{tag}_ncu_summary.md",
                   "workload": workload,
                   "sm__pipe_fp64_cycles_active_pct": pp.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                   "sm__pipe_tensor_cycles_active_pct": pp.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                   "smsp__issue_active_pct": pp.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   "sm__warps_active_pct": pp.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
                   "lts__t_sector_hit_rate_pct": pp.get("lts__t_sector_hit_rate.pct")},
                  open(os.path.join(out_dir, "qp_kernel_traffic.json"), "w"), indent=1)
    print(open(os.path.join(out_dir, f"{tag}_launches.md")).read())
    print(open(os.path.join(out_dir, f"{tag}_ncu_summary.md")).read()[:3000])


if __name__ == "__main__":
    main()

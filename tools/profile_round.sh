#!/bin/sh
# Round profile on the GPU box (run through gpurun):  sh tools/profile_round.sh <tag>
# 1. bench.py without a profiler (the numbers);  2. its launch list under ncu (shares of the step);
# 3. one `ncu --set full` capture of the two hot kernels of bench.py;  4. the same for the large-batch
# kernels (16,384 instances: throughput variant of qp_kernel, persistent qp8_kernel on both models).
# Raw outputs go to gpurun_out/; tools/summarize_profiles.py turns them into profiles/<tag>_*.md here.
set -e
TAG=${1:-r02}
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu --no-large --no-quad12 > gpurun_out/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'linearize_kernel|qp_kernel' -s 8 -c 2 -f \
    -o gpurun_out/${TAG}_full python bench.py --steps 6 --warmup 3 --no-cpu --no-large --no-quad12 > gpurun_out/${TAG}_ncu_full.log 2>&1
[ "$2" = bench ] && { ls -la gpurun_out; exit 0; }   # `sh tools/profile_round.sh <tag> bench`: only the bench step's kernels
# large batches: the one-instance kernel (forced: chunks of this size default to qp8_kernel), then qp8_kernel on both
# models.  gpurun brings back at most 64 MiB: of these captures only the raw-page CSV and the per-line profile travel.
LIB=mpc_blaster_b200/lib/libmpcb.so
cap() {  # cap <name> <kernel regex> <skip> <count> <sweep args...>
    NAME=$1; KR=$2; SK=$3; CN=$4; shift 4
    ncu --set full --clock-control none --import-source on -k regex:"$KR" -s $SK -c $CN -f -o gpurun_out/${TAG}_$NAME \
        python tools/sweep.py "$@" > gpurun_out/${TAG}_ncu_$NAME.log 2>&1
    ncu -i gpurun_out/${TAG}_$NAME.ncu-rep --page raw --csv > gpurun_out/${TAG}_${NAME}_raw.csv
}
cap qp1_blaster17_16k 'linearize_kernel|qp_kernel' 2 2 --points "16384,20,17,rand" --qp8-batch 1000000000
# (since the end of round 2 that is the latency variant: no default chunk size selects the single-buffer variant any more)
python tools/ncu_lines.py gpurun_out/${TAG}_qp1_blaster17_16k.ncu-rep $LIB qp_kernelILi17ELi6ELi1ELi2ELi1ELb0 > gpurun_out/${TAG}_qp1_latency_16k_kernel_lines.txt 2>&1 || true
rm -f gpurun_out/${TAG}_qp1_blaster17_16k.ncu-rep
cap qp8_blaster17_16k 'qp8_kernel' 1 1 --points "16384,20,17,rand"
python tools/ncu_lines.py gpurun_out/${TAG}_qp8_blaster17_16k.ncu-rep $LIB qp8_kernelILi17ELi6 > gpurun_out/${TAG}_qp8_kernel_lines.txt 2>&1 || true
rm -f gpurun_out/${TAG}_qp8_blaster17_16k.ncu-rep
cap qp8_quad12_16k 'linearize_kernel|qp8_kernel' 2 2 --points "16384,20,12,rand"
rm -f gpurun_out/${TAG}_qp8_quad12_16k.ncu-rep
ls -la gpurun_out

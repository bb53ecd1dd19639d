#!/bin/sh
# A/B of the factorisation experiments prepared at the end of round 1, in ONE gpurun call.
#   here (no GPU):   sh tools/ab_prepared.sh build
#   on the GPU box:  gpurun --timeout 420 -- 'sh tools/ab_prepared.sh run'
# Variants (all -D switches of mpcb_qp.cuh; `base` is the product):
#   lq          Householder LQ on every iteration                      (round-1 state before the hybrid)
#   gramx       + LQ input pivots / normal-equations state block while 1e-7 < mu <= 1e-4
#   gramx_only  the same on every iteration with mu > 1e-7 (no full-Gram phase: 15 KB less code)
#   gramwin     full-Gram phase with the 12-entries-per-lane window Gram
#   mu5         hybrid switch at mu > 1e-5 instead of 1e-4
# Output: gpurun_out/ab_prepared.txt (timings at 1,024 BLASTER17 / QUAD12 instances) and, per variant that is faster
# than `base`, the verdict of the GPU suite run against it (MPCB_LIB_OVERRIDE).
set -e
cd "$(dirname "$0")/.."
if [ "$1" = build ]; then
    python tools/ab.py build base= lq=MPCB_GRAM_MU=1e30 gramx=MPCB_GRAM_X=1e-7 gramx_only=MPCB_GRAM_X=1e-7,MPCB_GRAM_MU=1e30 \
        gramwin=MPCB_GRAM_WINDOW mu5=MPCB_GRAM_MU=1e-5
    exit 0
fi
mkdir -p gpurun_out
OUT=gpurun_out/ab_prepared.txt
python tools/ab.py run base lq gramx gramx_only gramwin mu5 --points "1024,20,17,rand;1024,20,12,rand" > $OUT 2>&1
cat $OUT
for v in gramx gramx_only gramwin mu5; do
    echo "== GPU suite with $v" >> $OUT
    MPCB_LIB_OVERRIDE=$PWD/mpc_blaster_b200/lib/libmpcb_$v.so timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $OUT
done
# crossover between the latency variant (now with the hybrid factorisation) and the single-buffer throughput variant
echo "== MPCB_THROUGHPUT_BATCH: default (4096) against latency variant forced" >> $OUT
python tools/ab.py run base --points "4096,20,17,rand;6144,20,17,rand" >> $OUT 2>&1
MPCB_THROUGHPUT_BATCH=100000000 python tools/ab.py run base --points "4096,20,17,rand;6144,20,17,rand" >> $OUT 2>&1
tail -30 $OUT

#!/bin/bash
# ncu --set full capture of one T-class k_encode launch per encoder organisation (run under gpurun).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "plain 0" "ring 2"; do
  set -- $v
  CMB200_ENC_MODE=$2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_encode -s 1 -c 1 \
    -f -o gpurun_out/r2_encT_$1 python tools/kernel_bench.py --chunks ${CH:-4096} --classes ${CLS:-T} --reps 1 > gpurun_out/r2_prof_$1.log 2>&1
  tail -2 gpurun_out/r2_prof_$1.log
done

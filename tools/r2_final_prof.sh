#!/bin/bash
# Round-2 profile evidence of bench.py's launches (run under gpurun, one GPU).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_launches_bench.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_encode -s 3 -c 1 -f -o gpurun_out/r2_encode_blend \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_encode_blend.log 2>&1
echo "encode capture rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_decode -s 2 -c 1 -f -o gpurun_out/r2_decode_blend \
  python tools/kernel_bench.py --chunks 8192 --classes B --reps 2 > gpurun_out/r2_decode_blend.log 2>&1
echo "decode capture rc=$?"

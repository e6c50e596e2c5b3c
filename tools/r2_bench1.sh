#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?"; tail -c 1500 gpurun_out/r2_bench_n1.err; head -c 6000 gpurun_out/r2_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "rc=$?"; tail -c 800 gpurun_out/r2_bench_ref.err; cat gpurun_out/r2_bench_ref.json

#!/bin/bash
# bench.py at N GPUs the way the driver launches it (run under gpurun --gpus N).
cd "$(dirname "$0")/.."
N=${1:-${N:-2}}        # positional argument or N=… in the environment
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "rc=$?"; tail -c 3000 gpurun_out/r2_bench_n$N.err; head -c 5000 gpurun_out/r2_bench_n$N.json

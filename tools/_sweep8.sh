mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_8gpu_r1.log 2>&1
tail -1 gpurun_out/bench_8gpu_r1.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['e2e']['synchronous_call']['value'], d['index'], d['parity_spot_check'], d['clocks'])" || tail -30 gpurun_out/bench_8gpu_r1.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/multigpu_check.py 2>&1 | grep "exchange\]" | sort | head -16

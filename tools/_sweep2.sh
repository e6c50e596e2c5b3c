mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_store.py -m gpu -q --timeout 600 -x -k "put_step or remote_index or async_put" 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/multigpu_check.py 2>&1 | grep -v "^\*\|^$\|OMP_NUM" | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_2gpu_r1.log 2>&1
tail -1 gpurun_out/bench_2gpu_r1.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['e2e']['synchronous_call']['value'], d['index'], d['parity_spot_check'])" || tail -20 gpurun_out/bench_2gpu_r1.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['e2e']['synchronous_call']['value'], d['roofline']['frac'], d['parity_spot_check'])"

mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -5
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_async.json; python -c "
import json; d=json.load(open('gpurun_out/bench_async.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e']['synchronous_call']['value'], d['roofline']['frac'], d['parity_spot_check'])"

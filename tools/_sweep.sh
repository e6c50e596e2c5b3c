timeout 600 python -m pytest tests -m gpu -q --timeout 400 2>&1 | tail -2
timeout 200 python tools/kernel_bench.py --classes RB --chunks 16384 --reps 2 2>&1 | grep -E '^[RTZMB] '
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches_per_step')}, d['roofline']['frac'])"

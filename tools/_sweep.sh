mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_encode -s 3 -c 1 -o gpurun_out/encode_r1_16k_c python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e')}, d['roofline']['frac'])"
CMB200_SEG_KB=0 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('staged', {k:d[k] for k in ('value','ms_per_step','e2e')}, d['roofline']['frac'])"

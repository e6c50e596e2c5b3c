timeout 600 python -m pytest tests -m gpu -q --timeout 240 -x 2>&1 | tail -3
timeout 120 python tools/kernel_bench.py --classes RTZMB --chunks 4096 --reps 2 2>&1 | grep -E '^[RTZMB] ' | cut -c1-60

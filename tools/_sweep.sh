timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -3
timeout 600 python tools/api_threads_bench.py --per-thread 512 2>&1 | grep -v "^{" | tail -8

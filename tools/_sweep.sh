mkdir -p gpurun_out
timeout 300 python tools/kernel_bench.py --chunks 8192 --classes TMRZB --reps 2 2>&1 | python -c "
import sys,json
for line in sys.stdin:
    try:
        c=line[0]; d=json.loads(line[2:]); print(c, round(d['encode_gibs'],1), round(d['decode_gibs'],1), end='  ')
    except Exception: print(line.strip()[:200])
print()"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -4

mkdir -p gpurun_out
run() { echo "== warps=$1 ctas=$2 carve=$3"; CMB200_ENC_WARPS=$1 CMB200_ENC_CTAS_PER_SM=$2 CMB200_ENC_CARVEOUT=$3 timeout 300 python tools/kernel_bench.py --chunks 8192 --classes TMB --reps 2 2>&1 | python -c "
import sys,json
for line in sys.stdin:
    try:
        c=line[0]; d=json.loads(line[2:]); print(c, round(d['encode_gibs'],1), end='  ')
    except Exception: print(line.strip()[:200])
print()"; }
run 7 2 -1
run 7 2 100
run 6 2 86
run 6 2 100
run 12 1 86
run 13 1 100
run 5 2 72
run 10 1 72
run 4 2 58
run 8 1 58

#!/bin/bash
# Round-2 A/B of the encode kernel organisations on one B200 (run under gpurun).
# Each variant runs in its own process (the selectors are read once per process).
# CMB200_ENC_MODE: 0 lean loop / L1, 2 lean loop / TMA ring (the round-1 loop, mode 3, and the 8-lane
# groups, mode 1, were measured with this script before they were removed from the tree).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CH=${CH:-8192}
OUT=gpurun_out/${OUT:-r2_variants.txt}
run() {  # name, env...
  name=$1; shift
  echo "=== $name" | tee -a $OUT
  env "$@" timeout 300 python tools/kernel_bench.py --chunks $CH --classes ${CLASSES:-TMRZB} --reps 3 2>&1 | tee -a $OUT
}
: > $OUT
for v in ${VARIANTS:-lean:0:0 lean_fpna:0:1 ring:2:0 ring_fpna:2:1}; do
  IFS=: read name mode fpna <<< "$v"
  run $name CMB200_ENC_MODE=$mode CMB200_FP_NOALLOC=$fpna
done

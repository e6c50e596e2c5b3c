#!/bin/bash
# Round-2 A/B of the encode kernel organisations on one B200 (run under gpurun).
# Each variant runs in its own process (the selectors are read once per process).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CH=${CH:-8192}
run() {  # name, env...
  name=$1; shift
  echo "=== $name" | tee -a gpurun_out/r2_variants.txt
  env "$@" timeout 300 python tools/kernel_bench.py --chunks $CH --classes ${CLASSES:-TMRZB} --reps 3 2>&1 | tee -a gpurun_out/r2_variants.txt
}
: > gpurun_out/r2_variants.txt
run plain_l1        CMB200_ENC_MODE=0 CMB200_FP_NOALLOC=0
run plain_fpna      CMB200_ENC_MODE=0 CMB200_FP_NOALLOC=1
run ring_l1         CMB200_ENC_MODE=2 CMB200_FP_NOALLOC=0
run ring_fpna       CMB200_ENC_MODE=2 CMB200_FP_NOALLOC=1
run ring_fpna_12w   CMB200_ENC_MODE=2 CMB200_FP_NOALLOC=1 CMB200_RING_WARPS=12

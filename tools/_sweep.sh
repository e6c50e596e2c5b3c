mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/memcheck_r1.log 2>&1; tail -5 gpurun_out/memcheck_r1.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/racecheck_r1.log 2>&1; tail -5 gpurun_out/racecheck_r1.log
CMB200_ENC_MODE=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/memcheck_groups_r1.log 2>&1; tail -3 gpurun_out/memcheck_groups_r1.log

mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_encode -s 3 -c 1 -o gpurun_out/encode_r1_16k python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log

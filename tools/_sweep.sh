mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 2500 gpurun_out/bench_r1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_r1.json 2>> gpurun_out/bench_r1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_encode -s 12 -c 1 -o gpurun_out/encode_r1 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/config_timings.py > gpurun_out/config_timings.log 2>&1; echo rc=$? >> gpurun_out/config_timings.log; tail -9 gpurun_out/config_timings.log | cut -c1-330
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r01.csv \
   python bench.py --steps 1 --warmup 3 --snapshots-per-gpu 131072 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kf_gram_tma -s 6 -c 2 -o gpurun_out/prof_gram_tma_r01 \
   python bench.py --steps 1 --warmup 3 --snapshots-per-gpu 131072 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_full_n1.json 2> gpurun_out/bench_full.err; echo "bench full rc=$?"
cut -c1-400 gpurun_out/bench_full_n1.json; tail -3 gpurun_out/bench_full.err

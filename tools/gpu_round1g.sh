#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r01.csv \
   python bench.py --steps 1 --warmup 3 --snapshots-per-gpu 131072 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kf_gram_tile -s 6 -c 2 -o gpurun_out/prof_gram_r01 \
   python bench.py --steps 1 --warmup 3 --snapshots-per-gpu 131072 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:kf_lift_level -s 4 -c 2 -o gpurun_out/prof_lift_r01 \
   python bench.py --steps 1 --warmup 3 --snapshots-per-gpu 131072 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_lift.log 2>&1
ls -la gpurun_out | tail -8

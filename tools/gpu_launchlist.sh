#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 1 --warmup 3 --snapshots-per-gpu 131072 --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_bench.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/launches_r01_final.csv

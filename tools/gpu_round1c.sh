#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 tools/_build/gemm_bench 1536 > gpurun_out/gemm_bench.txt 2>&1; cat gpurun_out/gemm_bench.txt
timeout 300 tools/_build/gemm_bench 3072 > gpurun_out/gemm_bench_3072.txt 2>&1; cat gpurun_out/gemm_bench_3072.txt
timeout 600 python bench.py --steps 2 --warmup 3 --snapshots-per-gpu 1048576 --no-cpu-baseline > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err; echo "bench rc=$?"
cat gpurun_out/bench_1m.json; tail -5 gpurun_out/bench_1m.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --snapshots-per-gpu 65536 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_full_n1.json 2> gpurun_out/bench_full.err; echo "bench full rc=$?"
cat gpurun_out/bench_full_n1.json; tail -3 gpurun_out/bench_full.err

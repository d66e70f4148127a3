#!/bin/bash
# first GPU round: peaks, smoke, gpu tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 tools/_build/dmma_peak > gpurun_out/dmma_peak.txt 2>&1
timeout 300 python tools/fp64_peak.py > gpurun_out/fp64_peak.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/dmma_peak.txt; cat gpurun_out/fp64_peak.log | tail -2; tail -5 gpurun_out/smoke.log; tail -30 gpurun_out/pytest_gpu.log

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/batch_timing.py > gpurun_out/batch_timing.log 2>&1; echo "batch rc=$?"
tail -3 gpurun_out/batch_timing.log | cut -c1-500

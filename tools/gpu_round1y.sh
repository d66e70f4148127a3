#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_qp.py -m gpu -x -q -k "active_set or config3a" > gpurun_out/pytest_as.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_as.log
tail -3 gpurun_out/pytest_as.log
KF_SWEEP_N=64 timeout 400 python tools/config3_sweep.py > gpurun_out/config3_sweep.log 2>&1; echo "sweep rc=$?"
tail -1 gpurun_out/config3_sweep.log | cut -c1-300

#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_qp.py -m gpu -x -q -k "active_set" > gpurun_out/pytest_as.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_as.log
tail -5 gpurun_out/pytest_as.log
KF_SWEEP_N=${KF_SWEEP_N:-44} timeout 900 python tools/config3_sweep.py > gpurun_out/config3_sweep.log 2>&1; echo "sweep rc=$?"
tail -50 gpurun_out/config3_sweep.log | cut -c1-250

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python tools/config_timings.py > gpurun_out/config_timings.log 2>&1; echo rc=$? >> gpurun_out/config_timings.log; tail -3 gpurun_out/config_timings.log | cut -c1-330
KF_QP_BUDGETS=0,21,42 KF_QP_SWEEPS=300 timeout 240 python tools/config3_check.py > gpurun_out/config3_check.log 2>&1; echo rc=$? >> gpurun_out/config3_check.log; tail -5 gpurun_out/config3_check.log | cut -c1-400
timeout 400 python tools/config5_parity.py 16384 > gpurun_out/config5_parity.log 2>&1; echo rc=$? >> gpurun_out/config5_parity.log; tail -28 gpurun_out/config5_parity.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2

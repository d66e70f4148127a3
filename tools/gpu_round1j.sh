#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 120 tools/_build/gemm_bench 4096 > gpurun_out/gemm_bench4.txt 2>&1; echo rc=$? >> gpurun_out/gemm_bench4.txt; tail -8 gpurun_out/gemm_bench4.txt
KF_QP_BUDGETS=0,21,42 KF_QP_SWEEPS=300 timeout 240 python tools/config3_check.py > gpurun_out/config3_check.log 2>&1; echo rc=$? >> gpurun_out/config3_check.log; tail -8 gpurun_out/config3_check.log | cut -c1-400

#!/bin/bash
# 2 GPUs: partition test on one GPU, then the column-split config-3a sweep under torchrun
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_qp.py -m gpu -x -q -k "active_set or config3a" > gpurun_out/pytest_as.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_as.log
tail -4 gpurun_out/pytest_as.log
N=${KF_NGPU:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/config3_sweep_sharded.py > gpurun_out/sweep_n$N.log 2>&1; echo "sweep rc=$?"
tail -5 gpurun_out/sweep_n$N.log | cut -c1-600

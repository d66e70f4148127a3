#!/bin/bash
mkdir -p gpurun_out
N=${KF_NGPU:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/config3_sweep_sharded.py > gpurun_out/sweep_n$N.log 2>&1; echo "sweep rc=$?"
tail -3 gpurun_out/sweep_n$N.log | cut -c1-600

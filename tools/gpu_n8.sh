#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
free -g | head -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 2 --warmup 3 --snapshots-per-gpu 4194304 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"
cut -c1-300 gpurun_out/bench_n8.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_n8.err | tail -5

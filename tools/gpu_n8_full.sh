#!/bin/bash
# the named job: 1e8 snapshots, n=12, m=3, P=4096 bilinear, snapshot-sharded over 8 B200 (1.25e7 per rank)
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 1 --warmup 1 --e2e-steps 1 --no-fast-mode --no-cpu-baseline > gpurun_out/bench_n8_full.json 2> gpurun_out/bench_n8_full.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n8_full.json'))
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'total snapshots', d['config']['total_snapshots'], 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']), 'clocks', d['clocks'])
PY
tail -2 gpurun_out/bench_n8_full.err

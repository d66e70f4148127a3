#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for g in 1 0; do
timeout 600 python bench.py --steps 3 --warmup 3 --snapshots-per-gpu 1048576 --no-cpu-baseline --e2e-steps 1 --no-fast-mode --option graphs=$g > gpurun_out/bench_1m_g$g.json 2> gpurun_out/bench_1m_g$g.err; echo "bench graphs=$g rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_1m_g$g.json'))
r=d['roofline']
print('graphs=$g value',round(d['value']),'ms/step',round(d['ms_per_step'],1),'solve_ms',round(r['solve_ms_per_step'],1),'launches',d['gpu_launches'])
PY
done

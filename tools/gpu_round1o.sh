#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 --snapshots-per-gpu 2097152 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_2m.json 2> gpurun_out/bench_2m.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_2m.json'))
r=d['roofline']
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],1),'gram TF',round(r['achieved'],2),'frac',round(r['frac'],3),'solve',round(r['solve_ms_per_step'],1), 'e2e', round(d['e2e']['value']), 'fast', d['fast_mode'])
PY
tail -3 gpurun_out/bench_2m.err

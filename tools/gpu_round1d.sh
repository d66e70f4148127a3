#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
for mb in 24 48 64; do
timeout 600 python bench.py --steps 2 --warmup 3 --snapshots-per-gpu 1048576 --no-cpu-baseline --e2e-steps 1 --option panel_mb=$mb > gpurun_out/bench_1m_mb$mb.json 2> gpurun_out/bench_1m.err; echo "bench mb=$mb rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_1m_mb$mb.json'))
r=d['roofline']
print('mb=$mb value',round(d['value']),'ms/step',round(d['ms_per_step'],1),'gram TF',round(r['achieved'],2),'frac',round(r['frac'],3),'gram ms',round(r['gram_kernel_ms_per_step'],1),'liftgram',round(r['lift_gram_ms_per_step'],1),'solve',round(r['solve_ms_per_step'],1))
PY
done
tail -3 gpurun_out/bench_1m.err

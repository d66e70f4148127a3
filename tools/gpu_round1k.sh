#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for tma in 1 0; do
timeout 600 python bench.py --steps 2 --warmup 3 --snapshots-per-gpu 2097152 --no-cpu-baseline --e2e-steps 1 --option tma=$tma > gpurun_out/bench_2m_tma$tma.json 2> gpurun_out/bench_2m.err; echo "bench tma=$tma rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_2m_tma$tma.json'))
r=d['roofline']
print('tma=$tma value',round(d['value']),'ms/step',round(d['ms_per_step'],1),'gram TF',round(r['achieved'],2),'frac',round(r['frac'],3),'gram ms',round(r['gram_kernel_ms_per_step'],1),'liftgram',round(r['lift_gram_ms_per_step'],1),'solve',round(r['solve_ms_per_step'],1), 'e2e', round(d['e2e']['value']))
PY
done
tail -3 gpurun_out/bench_2m.err

#!/bin/bash
# full GPU suite, a 2M-snapshot bench, and an ncu capture of the active-set Cholesky kernel (config 3a, late heavy steps)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 --snapshots-per-gpu 2097152 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_2m.json 2> gpurun_out/bench_2m.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_2m.json'))
r=d['roofline']
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],1),'gram TF',round(r['achieved'],2),'frac',round(r['frac'],3),'e2e', round(d['e2e']['value']))
PY
KF_SWEEP_N=64 timeout 900 ncu --set full --clock-control none --import-source on -k regex:kf_as_chol -s 480 -c 2 -o gpurun_out/prof_as_chol_r01 python tools/config3_sweep.py > gpurun_out/ncu_as.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_as.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# round-end check on a fresh box: the DEFAULT bench run (full GPU suite / smoke: tools/gpu_round1p.sh)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks_final.csv &
SMI=$!
T0=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
echo "bench wall seconds: $(( $(date +%s) - T0 ))"
kill $SMI
python - <<PY
import json
d=json.load(open('gpurun_out/bench_default.json'))
r=d['roofline']
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],1),'gram TF',round(r['achieved'],2),'frac',round(r['frac'],3),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'],'cpu',round(d['cpu_baseline']['value']),'lift_only',round(d['lift_only']['frac'],3),'fast',round(d['fast_mode']['value']))
print(d['clocks'])
PY
tail -2 gpurun_out/bench_default.err

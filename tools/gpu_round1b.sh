#!/bin/bash
# second GPU round: new tests, first bench line on a 1M-snapshot shard, ncu launch list + full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 --snapshots-per-gpu 1048576 > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err; echo "bench rc=$?"
cat gpurun_out/bench_1m.json; tail -5 gpurun_out/bench_1m.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --snapshots-per-gpu 65536 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kf_gram_tile -s 4 -c 2 -o gpurun_out/prof_gram \
   python bench.py --steps 1 --warmup 3 --snapshots-per-gpu 65536 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out

#!/bin/bash
# One parameterised GPU-box script (replaces the per-experiment scripts of round 1).  Usage, through gpurun:
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh tests'        full GPU suite (-m gpu), log in gpurun_out/pytest_gpu.log
#   gpurun --timeout 900  -- 'bash tools/gpu_check.sh bench [args]' default bench line -> gpurun_out/bench_default.json (+ clocks)
#   gpurun --timeout 900  -- 'bash tools/gpu_check.sh launches'     ncu launch list of a short bench -> gpurun_out/launches.csv
#   gpurun --timeout 600  -- 'bash tools/gpu_check.sh lift'         K1-alone bandwidth table -> gpurun_out/lift_bw.json
#   gpurun --timeout 600  -- 'bash tools/gpu_check.sh sweep'        config-3a 64-budget lasso sweep -> gpurun_out/config3_sweep.json
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_check.sh scale 8 [args]'   torchrun bench at N ranks
set -u
mkdir -p gpurun_out
what=${1:-tests}; shift || true
case "$what" in
  tests)    timeout 1400 python -m pytest tests -x -q -m gpu --timeout 400 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log ;;
  bench)    timeout 800 python bench.py "$@" > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_default.json ;;
  launches) timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
              python bench.py --steps 1 --warmup 1 --snapshots-per-gpu 1000000 --no-cpu-baseline --no-lasso3a --no-fast-mode > gpurun_out/launches_bench.log 2>&1; echo "ncu rc=$?" ;;
  lift)     timeout 500 python tools/lift_bw.py | tail -6 ;;
  sweep)    timeout 500 python tools/config3_sweep.py | tail -1 ;;
  scale)    n=${1:-2}; shift || true
            timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus "$n" "$@" \
              > "gpurun_out/bench_n$n.json" 2> "gpurun_out/bench_n$n.err"; echo "bench rc=$?"; tail -c 600 "gpurun_out/bench_n$n.json" ;;
  *) echo "unknown: $what"; exit 2 ;;
esac

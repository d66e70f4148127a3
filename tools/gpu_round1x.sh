#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kf_lift_tile -s 2 -c 1 -o gpurun_out/prof_lift_tile_poly_r01 -f python tools/lift_bw.py > gpurun_out/ncu_lift_tile.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kf_lift_tile -s 10 -c 1 -o gpurun_out/prof_lift_tile_c5_r01 -f python tools/lift_bw.py > gpurun_out/ncu_lift_tile2.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/prof_lift_tile_*

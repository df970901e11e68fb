#!/bin/bash
# ncu --set full capture of one extrema_warp_kernel launch (SoA, velocity) into gpurun_out/ (run under gpurun).
mkdir -p gpurun_out
CMD="python tools/bench_extrema.py 262144"
timeout 300 $CMD > gpurun_out/r02_extrema.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:extrema_warp -s 3 -c 1 -o gpurun_out/r02_extrema_full $CMD > gpurun_out/ncu_extrema.log 2>&1
tail -2 gpurun_out/ncu_extrema.log

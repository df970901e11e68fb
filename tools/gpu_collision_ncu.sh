#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/bench_collision.py"
timeout 300 $CMD > gpurun_out/r02_collision.log 2>&1; cat gpurun_out/r02_collision.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:collision_kernel -s 3 -c 1 -o gpurun_out/r02_collision_full $CMD > gpurun_out/ncu_collision.log 2>&1
tail -2 gpurun_out/ncu_collision.log

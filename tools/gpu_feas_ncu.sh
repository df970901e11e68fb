#!/bin/bash
# ncu --set full capture of one feasibility-sweep launch (262,144 trajectories x 1,000 samples) into gpurun_out/
mkdir -p gpurun_out
CMD="python tools/bench_sweep.py --layout aos --batch 262144"
timeout 300 $CMD > gpurun_out/r02_sweep_bench.log 2>&1; grep -E "feasib|eval_range" gpurun_out/r02_sweep_bench.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"eval_tm_kernel.*int.3>" -s 2 -c 1 -o gpurun_out/r02_feas_full $CMD > gpurun_out/ncu_feas.log 2>&1
tail -2 gpurun_out/ncu_feas.log

#!/bin/bash
mkdir -p gpurun_out
for ov in 0 1 2; do
echo "overlap $ov"
MTG_SOLVE_OVERLAP=$ov timeout 300 python tools/step_probe.py 2>&1 | grep -v two_launch | cut -c1-200
done > gpurun_out/r02_step_probe_overlap.log 2>&1
cat gpurun_out/r02_step_probe_overlap.log

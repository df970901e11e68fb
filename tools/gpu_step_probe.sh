#!/bin/bash
# per-step time of the fused solve + argmin for CTA sizes 128 / 64 / 32, with consecutive solves overlapping
mkdir -p gpurun_out
for blk in 128 64 32; do
echo "block $blk"
MTG_SOLVE_BLOCK=$blk MTG_PROBE_OVERLAP=1 timeout 300 python tools/step_probe.py 2>&1 | grep -v two_launch | cut -c1-200
done > gpurun_out/r02_step_probe_block.log 2>&1
cat gpurun_out/r02_step_probe_block.log

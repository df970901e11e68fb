#!/bin/bash
# A/B of two builds in ONE gpurun call (box-to-box variance is ~5 %): put them at
# mav_tube_trajectory_generation_b200/libmtg_cuda_{A,B}.so; capi honours MTG_CUDA_LIB.
for i in 1 2; do
for v in A B; do
  echo "variant $v"; MTG_CUDA_LIB=$PWD/mav_tube_trajectory_generation_b200/libmtg_cuda_$v.so python tools/bench_sweep.py --layout aos --batch 262144 2>&1 | grep -E "eval_range|feasibility\(pos" | cut -c1-160
done; done

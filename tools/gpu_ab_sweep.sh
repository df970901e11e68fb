#!/bin/bash
# A/B of several builds in ONE gpurun call (box-to-box variance is ~5 %): put them at
# mav_tube_trajectory_generation_b200/libmtg_cuda_<V>.so (tools/build_variant.sh); capi honours MTG_CUDA_LIB.
# usage: tools/gpu_ab_sweep.sh A B C ...
for i in 1 2; do
for v in "$@"; do
  echo "variant $v"; MTG_CUDA_LIB=$PWD/mav_tube_trajectory_generation_b200/libmtg_cuda_$v.so python tools/bench_sweep.py --layout aos --batch 262144 2>&1 | grep -E "eval_range|feasibility" | cut -c1-160
done; done

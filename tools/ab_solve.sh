#!/bin/bash
mkdir -p gpurun_out
for v in "" s160; do
  if [ -n "$v" ]; then export MTG_CUDA_LIB=mav_tube_trajectory_generation_b200/libmtg_cuda_$v.so; fi
  for r in 1 2; do
  python bench.py --steps 20 --warmup 3 --no-sweep --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['status_nonzero'])"
  done
done

#!/bin/bash
# fused vs two-launch step of the driver's bench (run under gpurun)
mkdir -p gpurun_out
for flag in "" "--two-launch-step"; do
  for r in 1 2; do
  python bench.py --steps 20 --warmup 3 --no-sweep --no-cpu-baseline $flag 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$flag', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['gpu_launches'], d['status_nonzero'], d['sweep_argmin'])"
  done
done
python bench.py --steps 500 --warmup 20 --no-sweep --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('s500', d['value'], d['ms_per_step'], d['roofline']['kernel_share_of_step'])"

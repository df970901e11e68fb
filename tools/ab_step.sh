#!/bin/bash
# the driver's bench step: fused + overlapped (default) vs ordinary launches vs the two-launch step (run under gpurun)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sweep_gpu.py tests/test_solve_gpu.py -q -m gpu -x > gpurun_out/one_test.log 2>&1; tail -5 gpurun_out/one_test.log | cut -c1-300
for flag in "" "--no-overlap" "--no-overlap --two-launch-step"; do
  for r in 1 2; do
  python bench.py --steps 20 --warmup 3 --no-sweep --no-cpu-baseline $flag 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$flag', d['value'], d['ms_per_step'], r['kernel_ms'], r['kernel_ms_isolated'], r['frac'], r['frac_isolated'], d['gpu_launches'], d['status_nonzero'], d['sweep_argmin'])"
  done
done
python bench.py --steps 500 --warmup 20 --no-sweep --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('s500', d['value'], d['ms_per_step'], d['roofline']['kernel_share_of_step'])"

#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_8.log 2>&1; tail -8 gpurun_out/r02_pytest_gpu_8.log | cut -c1-250
python tools/sweep_clock_probe.py 1000000 12 > gpurun_out/r02_sweep_probe_planned.log 2>&1; cut -c1-400 gpurun_out/r02_sweep_probe_planned.log
for mode in split fused; do
  if [ $mode = fused ]; then export MTG_SOLVE_FUSED=1; else unset MTG_SOLVE_FUSED; fi
  python bench.py --steps 200 --warmup 20 --no-sweep --no-cpu-baseline > gpurun_out/r02_bench_solve_$mode.json 2> gpurun_out/r02_bench_solve_$mode.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_solve_$mode.json')); print('$mode', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e_cost_only']['value'])"
done
unset MTG_SOLVE_FUSED
CMD="python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"solve_canonical|coeffs_from_free" -s 6 -c 2 -o gpurun_out/r02_solve_split_full $CMD > gpurun_out/r02_ncu_solve.log 2>&1; echo "ncu rc=$?"

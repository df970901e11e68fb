#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_eval_gpu.py tests/test_sweep_gpu.py tests/test_nl_objective_gpu.py tests/test_shim_gpu.py -q -m gpu > gpurun_out/r02_pytest_gpu_7.log 2>&1; tail -8 gpurun_out/r02_pytest_gpu_7.log | cut -c1-250
python tools/sweep_clock_probe.py 1000000 12 > gpurun_out/r02_sweep_probe_planned.log 2>&1; cut -c1-700 gpurun_out/r02_sweep_probe_planned.log
MTG_EVAL_FUSED=1 python tools/sweep_clock_probe.py 1000000 12 > gpurun_out/r02_sweep_probe_fused.log 2>&1; cut -c1-700 gpurun_out/r02_sweep_probe_fused.log
CMD2="python tools/bench_sweep.py --layout aos --batch 1000000 --reps 2"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_ -s 2 -c 2 -o gpurun_out/r02_sweep_planned_full $CMD2 > gpurun_out/r02_ncu_sweep_planned.log 2>&1; echo "ncu rc=$?"

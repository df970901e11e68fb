#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_solve_generic_gpu.py tests/test_extrema_gpu.py tests/test_shim_gpu.py tests/test_nl_objective_gpu.py -q -m gpu > gpurun_out/r02_pytest_gpu_11.log 2>&1; tail -6 gpurun_out/r02_pytest_gpu_11.log | cut -c1-250
python tools/bench_extrema.py > gpurun_out/r02_extrema.log 2>&1; cat gpurun_out/r02_extrema.log
python - <<'PY'
import sys, json, numpy as np, torch
sys.path.insert(0, '.')
import bench, mav_tube_trajectory_generation_b200 as m
ctx = m.Context(0)
print(json.dumps(bench.extras_section(ctx, 6454.6)["solve_generic"]))
PY

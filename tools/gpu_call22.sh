#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_extrema_gpu.py tests/test_nl_objective_gpu.py tests/test_shim_gpu.py -q -m gpu > gpurun_out/r02_pytest_gpu_22.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu_22.log | cut -c1-300
timeout 300 python tools/bench_extrema.py > gpurun_out/r02_extrema_22.log 2>&1; cat gpurun_out/r02_extrema_22.log

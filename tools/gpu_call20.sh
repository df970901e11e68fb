#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_extrema_gpu.py tests/test_nl_objective_gpu.py tests/test_shim_gpu.py -q -m gpu > gpurun_out/r02_pytest_gpu_20.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu_20.log | cut -c1-300
timeout 300 python tools/bench_extrema.py > gpurun_out/r02_extrema_20.log 2>&1; cat gpurun_out/r02_extrema_20.log
for v in w8 w6; do
MTG_CUDA_LIB=mav_tube_trajectory_generation_b200/libmtg_cuda_$v.so timeout 300 python tools/bench_extrema.py > gpurun_out/r02_extrema_20_$v.log 2>&1; cat gpurun_out/r02_extrema_20_$v.log
done

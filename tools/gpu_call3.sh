#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_extrema_gpu.py -q -x > gpurun_out/r02_pytest_extrema.log 2>&1; tail -5 gpurun_out/r02_pytest_extrema.log
python tools/bench_extrema.py > gpurun_out/r02_extrema.log 2>&1; cat gpurun_out/r02_extrema.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:extrema_warp -s 8 -c 1 -o gpurun_out/r02_extrema_full python tools/bench_extrema.py > gpurun_out/r02_ncu_extrema.log 2>&1; echo "ncu rc=$?"
python -m pytest tests -q -m gpu -x > gpurun_out/r02_pytest_gpu_3.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu_3.log

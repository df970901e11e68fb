#!/bin/bash
# Extrema kernel check (run under gpurun): its parity tests, the callers' tests, and the timing of both layouts.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_extrema_gpu.py tests/test_nl_objective_gpu.py tests/test_shim_gpu.py -q -m gpu > gpurun_out/extrema_check_pytest.log 2>&1; tail -5 gpurun_out/extrema_check_pytest.log | cut -c1-300
timeout 300 python tools/bench_extrema.py > gpurun_out/extrema_check_bench.log 2>&1; cat gpurun_out/extrema_check_bench.log

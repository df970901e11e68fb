#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sweep_gpu.py -q -m gpu -x > gpurun_out/one_test.log 2>&1; tail -12 gpurun_out/one_test.log | cut -c1-300
timeout 300 python tools/step_probe.py > gpurun_out/r02_step_probe.log 2>&1; cat gpurun_out/r02_step_probe.log | cut -c1-200

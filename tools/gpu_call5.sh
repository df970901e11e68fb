#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_5.log 2>&1; tail -12 gpurun_out/r02_pytest_gpu_5.log | cut -c1-250

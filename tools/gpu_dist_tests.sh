#!/bin/bash
# world-size-2 GPU tests (torchrun inside the test) + the NCCL argmin check; run under `gpurun --gpus 2`
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sweep_dist_gpu.py -q -m gpu > gpurun_out/r02_pytest_dist_n2.log 2>&1; tail -4 gpurun_out/r02_pytest_dist_n2.log | cut -c1-250

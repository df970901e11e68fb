#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_collision_gpu.py -q -m gpu > gpurun_out/collision_check_pytest.log 2>&1; tail -15 gpurun_out/collision_check_pytest.log | cut -c1-300
timeout 300 python tools/bench_collision.py > gpurun_out/r02_collision.log 2>&1; cat gpurun_out/r02_collision.log | cut -c1-300

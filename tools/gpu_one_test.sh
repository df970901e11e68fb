#!/bin/bash
# usage (under gpurun): bash tools/gpu_one_test.sh <pytest args...>
mkdir -p gpurun_out
timeout 900 python -m pytest "$@" -q -m gpu -x > gpurun_out/one_test.log 2>&1; tail -30 gpurun_out/one_test.log | cut -c1-300

#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_eval_gpu.py -q -m gpu > gpurun_out/r02_pytest_gpu_12.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_12.log | cut -c1-250
for rep in 1 2; do
python tools/sweep_clock_probe.py 1000000 12 > gpurun_out/r02_sweep_probe_phys_$rep.log 2>&1; head -1 gpurun_out/r02_sweep_probe_phys_$rep.log | cut -c1-330
MTG_CUDA_LIB=$PWD/mav_tube_trajectory_generation_b200/libmtg_cuda_logical.so python tools/sweep_clock_probe.py 1000000 12 > gpurun_out/r02_sweep_probe_logical_$rep.log 2>&1; head -1 gpurun_out/r02_sweep_probe_logical_$rep.log | cut -c1-330
done
python tools/bench_sweep.py --layout aos --batch 1000000 --reps 5 2>&1 | cut -c1-250
MTG_CUDA_LIB=$PWD/mav_tube_trajectory_generation_b200/libmtg_cuda_logical.so python tools/bench_sweep.py --layout aos --batch 1000000 --reps 5 2>&1 | cut -c1-250

#!/bin/bash
# r02 call 1: regression of the whole -m gpu suite on the new table management, a driver-style bench line,
# and an ncu --set full capture of the position sweep at the size the bench times (1M trajectories).
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/r02_pytest_gpu_1.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_1.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1_s20.json 2> gpurun_out/r02_bench_n1_s20.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/r02_bench_n1_s20.json
CMD2="python tools/bench_sweep.py --layout aos --batch 1000000 --reps 1"
timeout 300 $CMD2 > gpurun_out/r02_plain_sweep_1m.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_tm -s 1 -c 1 -o gpurun_out/r02_sweep_full_1m $CMD2 > gpurun_out/r02_ncu_sweep_1m.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/r02_plain_sweep_1m.log

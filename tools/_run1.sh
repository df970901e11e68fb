set -x
python -m pytest tests/test_eval_gpu.py -x -q -m gpu > gpurun_out/pytest_eval_tm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_eval_tm.log
python tools/bench_sweep.py --layout aos --batch 262144 > gpurun_out/sweep_tm_aos.log 2>&1
python tools/bench_sweep.py --layout aos --batch 131072 --reps 2 > gpurun_out/plain_tm.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:eval_tm -c 2 -o gpurun_out/prof_eval_tm_v13 python tools/bench_sweep.py --layout aos --batch 131072 --reps 2 > gpurun_out/ncu_tm.log 2>&1
tail -3 gpurun_out/pytest_eval_tm.log; cat gpurun_out/sweep_tm_aos.log

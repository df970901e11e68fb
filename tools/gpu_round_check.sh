#!/bin/bash
# Round-end GPU check (run under gpurun from the repo root): full -m gpu suite, smoke, bench (both arms),
# ncu launch list and --set full captures of the solve / sweep / feasibility kernels into gpurun_out/.
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r1.log 2>&1; tail -3 gpurun_out/pytest_gpu_r1.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1.log 2>&1; tail -1 gpurun_out/smoke_r1.log
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1_reference.json 2>/dev/null; echo "ref rc=$?"
CMD="python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu-baseline"
$CMD > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain_bench2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:solve_canonical -s 3 -c 2 -o gpurun_out/r01_solve_full $CMD > gpurun_out/ncu_solve.log 2>&1
CMD2="python tools/bench_sweep.py --layout aos --batch 131072 --reps 2"
$CMD2 > gpurun_out/plain_sweep.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:eval_tm -s 1 -c 1 -o gpurun_out/r01_sweep_full $CMD2 > gpurun_out/ncu_sweep.log 2>&1
$CMD2 > gpurun_out/plain_sweep2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:eval_tm -s 4 -c 1 -o gpurun_out/r01_feas_full $CMD2 > gpurun_out/ncu_feas.log 2>&1
cut -c1-300 gpurun_out/bench_r1.json

#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_extrema_gpu.py tests/test_nl_objective_gpu.py -q -m gpu > gpurun_out/r02_pytest_gpu_10.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_10.log | cut -c1-250
python tools/bench_extrema.py > gpurun_out/r02_extrema.log 2>&1; cat gpurun_out/r02_extrema.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1_s20.json 2> gpurun_out/r02_bench_n1_s20.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1_s20.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1_s20.json')); print(d['value'], d['ms_per_step']); print(json.dumps(d['extras'])[:2500])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:extrema_warp -s 8 -c 1 -o gpurun_out/r02_extrema_full python tools/bench_extrema.py > gpurun_out/r02_ncu_extrema.log 2>&1; echo "ncu rc=$?"

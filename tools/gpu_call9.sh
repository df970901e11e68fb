#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_9.log 2>&1; tail -8 gpurun_out/r02_pytest_gpu_9.log | cut -c1-250
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1_s20.json 2> gpurun_out/r02_bench_n1_s20.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1_s20.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac']); print(json.dumps(d['sweep'])[:1400]); print(json.dumps(d['time_alloc'])[:500])"

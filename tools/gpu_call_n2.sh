#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_sweep_dist_gpu.py -q -m gpu > gpurun_out/r02_pytest_dist_n2.log 2>&1; tail -4 gpurun_out/r02_pytest_dist_n2.log | cut -c1-250
python bench.py --gpus 1 --steps 20 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/r02_bench_n1_for_scale.json 2>/dev/null
for rep in 1 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$rep bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_n2_s20_$rep.json 2> gpurun_out/r02_bench_n2_s20_$rep.err; echo "bench n2 rc=$?"
done
python -c "
import json
a=json.load(open('gpurun_out/r02_bench_n1_for_scale.json'))
print('n1', a['value'], a['ms_per_step'], a['ranks'])
for rep in (1,2):
    d=json.load(open('gpurun_out/r02_bench_n2_s20_%d.json'%rep))
    print('n2', d['value'], d['ms_per_step'], 'eff', d['value']/(2*a['value']), d['ranks'], d['e2e']['value'], d['sweep_argmin'])"

#!/bin/bash
# 1 -> 2 -> 4 -> 8 GPU scaling of the driver's bench (run under `gpurun --gpus 8`): each N launched the way the driver
# launches it, --steps 20 --warmup 3; the efficiency is value(N) / (N * value(1)).
R=${1:-r02}
NS=${2:-"2 4 8"}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/${R}_scale_n1.json 2>/dev/null; echo "n1 rc=$?"
port=29540
for n in $NS; do
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/${R}_scale_n$n.json 2> gpurun_out/${R}_scale_n$n.err; echo "n$n rc=$?"
done
python -c "
import json
a=json.load(open('gpurun_out/${R}_scale_n1.json'))
print('n1', a['value'], a['ms_per_step'])
for n in (2,4,8):
    try:
        d=json.load(open('gpurun_out/${R}_scale_n%d.json'%n))
        print('n%d'%n, d['value'], d['ms_per_step'], 'eff', round(d['value']/(n*a['value']),3), 'e2e', d['e2e']['value'], [round(x,3) for x in d['ranks']['ms_compute']], d['ranks']['comm_nranks_ok'])
    except Exception as e: print(n, 'failed', e)"

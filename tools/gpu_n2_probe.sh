#!/bin/bash
mkdir -p gpurun_out
port=29600
for flag in "" "--no-overlap" "--no-overlap --two-launch-step" ""; do
  port=$((port+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 20 --warmup 3 --no-sweep --no-cpu-baseline $flag 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('[$flag]', d['value'], d['ms_per_step'], [round(x,3) for x in d['ranks']['ms_compute']], [round(x,3) for x in d['ranks']['ms_total']])"
done

#!/bin/bash
# configs[4] (1 M-candidate sweep): 1 GPU with the oracle check, then N = 2 / 4 / 8 under torchrun (run under gpurun --gpus 8)
mkdir -p gpurun_out
timeout 600 python tools/sweep_config5.py 2>/dev/null | grep '^{' > gpurun_out/r02_config5_n1.json; cut -c1-600 gpurun_out/r02_config5_n1.json
port=29640
for n in ${1:-2 4 8}; do
  [ "$n" = "none" ] && break
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port tools/sweep_config5.py --no-oracle 2>/dev/null | grep '^{' > gpurun_out/r02_config5_n$n.json; cut -c1-420 gpurun_out/r02_config5_n$n.json
done

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 tools/pcie_probe.py > gpurun_out/r02_pcie_probe_n8.json 2> gpurun_out/pcie_probe.err; echo rc=$?; cat gpurun_out/r02_pcie_probe_n8.json; tail -3 gpurun_out/pcie_probe.err
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1; head -14 gpurun_out/r02_topo.txt | cut -c1-200

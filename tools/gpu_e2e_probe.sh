#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/e2e_probe.py > gpurun_out/r02_e2e_probe.log 2>&1; cat gpurun_out/r02_e2e_probe.log | cut -c1-250

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/step_probe.py > gpurun_out/r02_step_probe.log 2>&1; cat gpurun_out/r02_step_probe.log | cut -c1-200

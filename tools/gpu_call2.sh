#!/bin/bash
# r02 call 2: the warp-cooperative extrema kernel: memcheck on a small case, the parity suite, timing, ncu.
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_extrema_gpu.py -q -x -k "random_batch or real_roots or one_dimension" > gpurun_out/r02_extrema_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r02_extrema_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_extrema_gpu.py -q -x -k "random_batch" > gpurun_out/r02_extrema_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/r02_extrema_racecheck.log
python -m pytest tests/test_extrema_gpu.py -q > gpurun_out/r02_pytest_extrema.log 2>&1; tail -15 gpurun_out/r02_pytest_extrema.log
python tools/bench_extrema.py > gpurun_out/r02_extrema.log 2>&1; cat gpurun_out/r02_extrema.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:extrema_warp -s 2 -c 1 -o gpurun_out/r02_extrema_full python tools/bench_extrema.py > gpurun_out/r02_ncu_extrema.log 2>&1; echo "ncu rc=$?"

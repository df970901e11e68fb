#!/bin/bash
# the -m gpu suite N times in a row (flakiness check of the lock / work-stealing code paths)
mkdir -p gpurun_out
for i in $(seq 1 ${1:-3}); do
  timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/repeat_$i.log 2>&1; tail -1 gpurun_out/repeat_$i.log
done

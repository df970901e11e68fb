#!/bin/bash
# Compiles the C++ class-API check against the C ABI (no GPU needed to compile; running needs one).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
g++ -std=c++14 -O1 -Wall -Wextra -I "$ROOT/include" "$HERE/shim_check.cpp" \
    -L "$ROOT/mav_tube_trajectory_generation_b200" -lmtg_cuda \
    -Wl,-rpath,"$ROOT/mav_tube_trajectory_generation_b200" -Wl,-rpath,/usr/local/cuda/lib64 \
    -o "$HERE/shim_check"
g++ -std=c++14 -O1 -Wall -Wextra -I "$ROOT/include" -I /usr/local/cuda/include "$HERE/sweep_check.cpp" \
    -L "$ROOT/mav_tube_trajectory_generation_b200" -lmtg_cuda -L /usr/local/cuda/lib64 -lcudart \
    -Wl,-rpath,"$ROOT/mav_tube_trajectory_generation_b200" -Wl,-rpath,/usr/local/cuda/lib64 \
    -o "$HERE/sweep_check"
echo "built $HERE/shim_check $HERE/sweep_check"

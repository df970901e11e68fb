#!/bin/bash
# Builds mav_tube_trajectory_generation_b200/libmtg_cuda_<NAME>.so with extra -D flags applied to ONE
# translation unit (default eval_tm.cu), re-using the other objects of the regular build.
# usage: tools/build_variant.sh NAME "-DMTG_TM_BULK=0 ..." [unit]
set -e
NAME=$1; FLAGS=$2; UNIT=${3:-eval_tm}
PKG=$(dirname "$0")/../mav_tube_trajectory_generation_b200
B=$PKG/build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --fmad=true $FLAGS \
  -c $PKG/csrc/$UNIT.cu -o /tmp/${UNIT}_$NAME.o
OBJS=""
for o in $B/*.o; do [ "$(basename $o)" = "$UNIT.o" ] && OBJS="$OBJS /tmp/${UNIT}_$NAME.o" || OBJS="$OBJS $o"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $PKG/libmtg_cuda_$NAME.so $OBJS -ldl
echo built $PKG/libmtg_cuda_$NAME.so

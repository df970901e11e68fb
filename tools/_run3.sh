CMD="python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu-baseline"
$CMD > gpurun_out/plain_bench2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:solve_canonical -s 3 -c 1 -o gpurun_out/solve_v2 $CMD > gpurun_out/ncu_solve.log 2>&1
echo done

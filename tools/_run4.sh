CMD2="python tools/bench_sweep.py --layout aos --batch 131072 --reps 2"
$CMD2 > gpurun_out/plain_sweep.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:eval_tm -s 4 -c 1 -o gpurun_out/feas_v1 $CMD2 > gpurun_out/ncu_feas.log 2>&1
echo done

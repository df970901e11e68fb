#!/bin/bash
# Round-end GPU check (run under gpurun from the repo root): full -m gpu suite, smoke, bench (both arms, driver's
# and default step counts), ncu launch list and --set full captures of the solve and extrema kernels into gpurun_out/.
R=r02
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/${R}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${R}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${R}_bench_reference_n1.json 2>/dev/null; echo "ref rc=$?"
python bench.py --steps 20 --warmup 3 > gpurun_out/${R}_bench_n1_s20.json 2> gpurun_out/${R}_bench_n1_s20.err; echo "bench s20 rc=$?"
python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc=$?"
CMD="python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu-baseline"
$CMD > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain_bench2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:solve_canonical -s 3 -c 2 -o gpurun_out/${R}_solve_full $CMD > gpurun_out/ncu_solve.log 2>&1
ECMD="python tools/bench_extrema.py 262144"
$ECMD > gpurun_out/${R}_extrema.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:extrema_warp -s 3 -c 1 -o gpurun_out/${R}_extrema_full $ECMD > gpurun_out/ncu_extrema.log 2>&1
python -c "
import json
for f in ('s20',''):
    d=json.load(open('gpurun_out/${R}_bench_n1'+('_'+f if f else '')+'.json')); print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['sweep']['eval_range']['frac'], d['sweep']['feasibility']['frac'], d['sweep']['extrema_v_and_a']['value'], d['sweep']['parity_sample']['ok'])
r=json.load(open('gpurun_out/${R}_bench_reference_n1.json')); print('ref', r['value'], r['cpu_baseline']['cores'])"

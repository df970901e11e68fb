#!/usr/bin/env python
"""Why do CUDA events around the position sweep (bench: 6.5-6.6 ms at 1M trajectories) disagree with the ncu
duration of the same launch (5.46 ms, profiles/r02_sweep_full_1m.txt)? Times N back-to-back launches one by
one (an event pair per launch) while NVML samples the SM / memory clocks and the power draw: if the first
launches after an idle gap are slow and the clocks are still ramping, the difference is warm-up, not the kernel."""
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mav_tube_trajectory_generation_b200 as m  # noqa: E402
import pynvml  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
ctx = m.Context(0)
pos, times = bench.make_workload(B, 4)
pos = np.ascontiguousarray(np.moveaxis(pos, -1, 0))
times = np.ascontiguousarray(np.moveaxis(times, -1, 0))
p, t = torch.from_numpy(pos).cuda(), torch.from_numpy(times).cuda()
sol = ctx.solve_batch(p, t, layout="aos")
tmax = ctx.max_time_batch(t, layout="aos")
dt = tmax / 1000
samples = torch.empty((B, 1008, 3), dtype=torch.float64, device="cuda")
r = ctx.eval_range_batch(sol["coeffs"], t, 0.0, tmax, dt, 0, 1008, layout="aos", out={"samples": samples})
nsamp = int(r["n_samples"].sum().item())
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
log, stop = [], threading.Event()


def poll():
    while not stop.is_set():
        log.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                    pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3,
                    int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))))
        time.sleep(0.001)


for idle in (0.0, 2.0):
    torch.cuda.synchronize()
    time.sleep(idle)          # an idle gap before the burst, like the one in front of bench.py's sweep section
    log.clear()
    stop.clear()
    th = threading.Thread(target=poll, daemon=True)
    th.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        ctx.eval_range_batch(sol["coeffs"], t, 0.0, tmax, dt, 0, 1008, layout="aos",
                             out={"samples": samples, "n_samples": r["n_samples"]})
        ev[i + 1].record()
    torch.cuda.synchronize()
    stop.set()
    th.join()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
    sm = [x[1] for x in log]
    print(json.dumps({"idle_before_s": idle, "batch": B, "samples": nsamp, "ms_per_launch": [round(x, 3) for x in ms],
                      "frac_of_hbm_best": 26.48 * nsamp / (min(ms) * 1e-3) / 1e9 / bench.peaks()[0],
                      "frac_of_hbm_median": 26.48 * nsamp / (float(np.median(ms)) * 1e-3) / 1e9 / bench.peaks()[0],
                      "sm_mhz_min_med_max": [min(sm), float(np.median(sm)), max(sm)],
                      "mem_mhz_min_max": [min(x[2] for x in log), max(x[2] for x in log)],
                      "power_w_max": max(x[3] for x in log),
                      "reasons_or": hex(int(np.bitwise_or.reduce([x[4] for x in log])))}), flush=True)

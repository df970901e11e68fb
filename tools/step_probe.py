#!/usr/bin/env python
"""Per-launch time of the solve alone, the fused solve + argmin, and the two-launch step, same rotating buffers."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import mav_tube_trajectory_generation_b200 as m

ctx = m.Context(0)
if os.environ.get("MTG_PROBE_OVERLAP"):
    ctx.set_solve_overlap(True)   # inputs are resident: consecutive solves may overlap
B, R = 65536, 4
ins, outs = [], []
for i in range(R):
    pos, times = bench.make_workload(B, seed=i)
    ins.append((torch.from_numpy(pos).cuda(), torch.from_numpy(times).cuda()))
    outs.append({"coeffs": torch.empty((10, 3, 10, B), dtype=torch.float64, device="cuda"),
                 "cost": torch.empty((B,), dtype=torch.float64, device="cuda"),
                 "status": torch.empty((B,), dtype=torch.int32, device="cuda")})
best = torch.zeros(2, dtype=torch.int64, device="cuda")


def solve(i):
    ctx.solve_batch(*ins[i % R], out=outs[i % R])


def fused(i):
    ctx.solve_argmin_batch(*ins[i % R], out=outs[i % R], global_offset=i * B, best=best, accumulate=i > 0)


def two(i):
    o = outs[i % R]
    ctx.solve_batch(*ins[i % R], out=o)
    ctx.argmin_batch(o["cost"], status=o["status"], global_offset=i * B, best=best, accumulate=i > 0)


for name, fn in (("solve", solve), ("fused", fused), ("two_launch", two), ("solve", solve), ("fused", fused)):
    for n in (20, 200):
        for i in range(5):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"step": name, "launches": n, "us_per_step": e0.elapsed_time(e1) / n * 1e3}))

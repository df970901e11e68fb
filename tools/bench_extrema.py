#!/usr/bin/env python
"""Extrema kernel timing (root problems/s) for both layouts; prints JSON lines."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import mav_tube_trajectory_generation_b200 as m

ctx = m.Context(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
pos, times = bench.make_workload(B, 4)
for layout in ("soa", "aos"):
    p, t = pos, times
    if layout == "aos":
        p, t = np.ascontiguousarray(np.moveaxis(pos, -1, 0)), np.ascontiguousarray(np.moveaxis(times, -1, 0))
    p, t = torch.from_numpy(p).cuda(), torch.from_numpy(t).cuda()
    sol = ctx.solve_batch(p, t, layout=layout)
    for der in (1, 2):
        ctx.extrema_batch(sol["coeffs"], t, der, layout=layout)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = ctx.extrema_batch(sol["coeffs"], t, der, layout=layout)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(json.dumps({"layout": layout, "derivative": der, "batch": B, "ms": ms,
                          "root_problems_per_s": B * 10 / ms * 1e3, "status_nonzero": int((r["status"] != 0).sum())}))

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
        for _ in range(3):
            ctx.extrema_batch(sol["coeffs"], t, der, layout=layout)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record()
        for i in range(5):
            r = ctx.extrema_batch(sol["coeffs"], t, der, layout=layout)
            ev[i + 1].record()
        torch.cuda.synchronize()
        all_ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(5))
        ms = all_ms[2]
        print(json.dumps({"layout": layout, "derivative": der, "batch": B, "ms": ms, "ms_best": all_ms[0],
                          "root_problems_per_s": B * 10 / ms * 1e3, "status_nonzero": int((r["status"] != 0).sum())}))

#!/usr/bin/env python
"""End-to-end (pinned host memory) solve rate for both layouts and a few staging chunk sizes."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import mav_tube_trajectory_generation_b200 as m

B = 65536
pos, times = bench.make_workload(B, 4)
for chunk in (None, 4096, 16384):
    if chunk is None:
        os.environ.pop("MTG_HOST_CHUNK", None)
    else:
        os.environ["MTG_HOST_CHUNK"] = str(chunk)
    ctx = m.Context(0)
    for layout in ("soa", "aos"):
        p, t = pos, times
        shape = (10, 3, 10, B)
        if layout == "aos":
            p, t = np.ascontiguousarray(np.moveaxis(pos, -1, 0)), np.ascontiguousarray(np.moveaxis(times, -1, 0))
            shape = (B, 10, 3, 10)
        hp, ht = torch.from_numpy(p).pin_memory(), torch.from_numpy(t).pin_memory()
        out = {"coeffs": torch.empty(shape, dtype=torch.float64).pin_memory(),
               "cost": torch.empty((B,), dtype=torch.float64).pin_memory(),
               "status": torch.empty((B,), dtype=torch.int32).pin_memory()}
        for _ in range(3):
            ctx.solve_batch(hp, ht, N=10, derivative=4, out=out, layout=layout)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            ctx.solve_batch(hp, ht, N=10, derivative=4, out=out, layout=layout)
            _ = float(out["cost"][0])
        torch.cuda.synchronize()
        s = (time.perf_counter() - t0) / 10
        print(json.dumps({"chunk": chunk or "default (8192)", "layout": layout, "ms": s * 1e3, "trajectories_per_s": B / s,
                          "d2h_gbs": B * 2412 / s / 1e9}))

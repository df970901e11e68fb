#!/usr/bin/env python
"""Collision potential (N4) timing: 65,536 trajectories of the configs[1] shape against a dense distance grid."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import mav_tube_trajectory_generation_b200 as m

ctx = m.Context(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
pos, times = bench.make_workload(B, 4)
p, t = torch.from_numpy(pos).cuda(), torch.from_numpy(times).cuda()
sol = ctx.solve_batch(p, t, layout="soa")
g = torch.full((120, 120, 120), 3.0, dtype=torch.float64, device="cuda")
for grad in (True, False):
    fn = lambda: ctx.collision_cost_batch(sol["coeffs"], t, g, [-60, -60, -60], 0.2, [-40.0] * 3, [40.0] * 3, 0.1,
                                          epsilon=4.0, robot_radius=0.3, want_grad=grad, layout="soa")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    for i in range(5):
        r = fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(5))[2]
    print(json.dumps({"gradient": grad, "batch": B, "ms": ms, "trajectories_per_s": B / ms * 1e3,
                      "checks_mean": float(r["n_checks"].double().mean()),
                      "in_collision": int(r["in_collision"].sum())}))

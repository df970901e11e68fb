#!/usr/bin/env python
"""BASELINE config 4: dense evaluateRange + v/a/tube feasibility sweep, S ~ 1000 samples per
trajectory, device-resident. Prints one JSON line per kernel variant (samples/s, achieved
GB/s against the measured HBM peak). Not the driver's bench: a measurement tool for DESIGN.md."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mav_tube_trajectory_generation_b200 as m  # noqa: E402


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=262144)
    ap.add_argument("--samples", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--layout", default="soa")
    args = ap.parse_args()
    B, S = args.batch, args.samples
    ctx = m.Context(0)
    pos, times = bench.make_workload(B, 4)
    if args.layout == "aos":
        pos = np.ascontiguousarray(np.moveaxis(pos, -1, 0))
        times = np.ascontiguousarray(np.moveaxis(times, -1, 0))
    p, t = torch.from_numpy(pos).cuda(), torch.from_numpy(times).cuda()
    sol = ctx.solve_batch(p, t, layout=args.layout)
    coeffs = sol["coeffs"]
    secs = timeit(lambda: ctx.solve_batch(p, t, layout=args.layout, out=sol), args.reps)
    print(json.dumps({"kernel": "solve_canonical", "layout": args.layout, "batch": B, "ms": secs * 1e3,
                      "trajectories_per_s": B / secs, "achieved_gbs": bench.BYTES_PER_TRAJ * B / secs / 1e9}), flush=True)
    tmax = ctx.max_time_batch(t, layout=args.layout)
    dt = tmax / S
    Smax = S + 8
    shape_s = (B, Smax, 3) if args.layout == "aos" else (Smax, 3, B)
    shape_f = (B, Smax) if args.layout == "aos" else (Smax, B)
    samples = torch.empty(shape_s, dtype=torch.float64, device="cuda")
    flags = torch.empty(shape_f, dtype=torch.uint8, device="cuda")
    radii = torch.full((B, 10, 2) if args.layout == "aos" else (10, 2, B), 0.15, dtype=torch.float64, device="cuda")
    peak, src = bench.peaks()
    n = None

    def report(name, secs, bytes_per_sample, nsamp):
        gbs = bytes_per_sample * nsamp / secs / 1e9
        print(json.dumps({"kernel": name, "layout": args.layout, "batch": B, "samples_per_traj": S,
                          "samples_per_s": nsamp / secs, "ms": secs * 1e3, "achieved_gbs": gbs,
                          "bytes_per_sample": bytes_per_sample, "hbm_peak": peak, "frac": gbs / peak}), flush=True)

    r = ctx.eval_range_batch(coeffs, t, 0.0, tmax, dt, 0, Smax, layout=args.layout, out={"samples": samples})
    torch.cuda.synchronize()
    nsamp = int(r["n_samples"].sum().item())
    assert int(r["status"].max().item()) == 0
    secs = timeit(lambda: ctx.eval_range_batch(coeffs, t, 0.0, tmax, dt, 0, Smax, layout=args.layout,
                                               out={"samples": samples, "n_samples": r["n_samples"]}), args.reps)
    report("eval_range(position)", secs, 24 + 2.48, nsamp)
    secs = timeit(lambda: ctx.feasibility_batch(coeffs, t, 0.0, tmax, dt, 3.0, 5.0, positions=p, radii=radii,
                                                max_samples=Smax, layout=args.layout, want_samples=True,
                                                out={"samples": samples, "flags": flags}), args.reps)
    report("feasibility(pos+flags+tube)", secs, 25 + 2.48, nsamp)
    secs = timeit(lambda: ctx.feasibility_batch(coeffs, t, 0.0, tmax, dt, 3.0, 5.0, positions=p, radii=radii,
                                                max_samples=Smax, layout=args.layout, want_samples=False,
                                                out={"flags": flags}), args.reps)
    report("feasibility(flags+tube only)", secs, 1 + 2.48, nsamp)


if __name__ == "__main__":
    main()

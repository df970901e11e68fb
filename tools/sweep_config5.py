#!/usr/bin/env python
"""BASELINE configs[4] / SURVEY config 5: a 1,000,000-candidate random-vertex min-snap sweep sharded contiguously
over the ranks (one process per GPU under torchrun, or a single GPU), NCCL used only for the final argmin gather.

Every rank generates ITS shard on the device (Philox keyed by the global candidate index, so the candidates do not
depend on the split), solves it with the argmin fused into the solve kernel (mtg_solve_argmin_batch, one launch), and
the {cost, index} pairs are all-gathered and folded on the device (mtg_best_allgather on the library's own
communicator). Timed with CUDA events, max over ranks, 3 warm-up sweeps. Rank 0 then solves ALL candidates with the
CPU oracle (test infrastructure, all host threads) and checks the global pair against its serial scan.
Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import mav_tube_trajectory_generation_b200 as m
from mav_tube_trajectory_generation_b200 import sweep

TOTAL, K, D, SEED = 1_000_000, 10, 3, 20261019
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = m.Context(local)
if world > 1:
    uid = [ctx.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.nccl_init(uid[0], rank, world)
start, count = sweep.shard_range(TOTAL, rank, world)
best = torch.zeros(2, dtype=torch.int64, device="cuda")
best_global = torch.zeros(2, dtype=torch.int64, device="cuda")


def generate():
    return ctx.generate_candidates_batch(count, K, D, seed=SEED, first_index=start, pos_min=[-10.0] * D,
                                         pos_max=[10.0] * D, v_max=3.0, a_max=5.0)


def run(pos, times):
    ctx.solve_argmin_batch(pos, times, global_offset=start, best=best, accumulate=False)
    ctx.best_allgather(best, out=best_global)


pos, times = generate()
for _ in range(3):
    run(pos, times)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
reps = 10
ev[0].record()
for _ in range(reps):
    pos, times = generate()
ev[1].record()
for _ in range(reps):
    run(pos, times)
ev[2].record()
torch.cuda.synchronize()
t = torch.tensor([ev[0].elapsed_time(ev[1]) / reps, ev[1].elapsed_time(ev[2]) / reps], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
gen_ms, sweep_ms = float(t[0]), float(t[1])
cost, idx = ctx.decode_best(best_global)
line = {"workload": "configs[4]: 1,000,000 random 3-D 10-segment min-snap candidates, contiguous shards, one "
                    "mtg_solve_argmin_batch per rank + mtg_best_allgather",
        "n_gpus": world, "candidates": TOTAL, "per_gpu": count, "sweep_ms": sweep_ms,
        "value": TOTAL / (sweep_ms * 1e-3), "unit": "candidates solved/s (solve + argmin + gather, device resident)",
        "generate_ms": gen_ms, "argmin": {"cost": cost, "candidate": idx}}
if rank == 0 and "--no-oracle" not in sys.argv:
    # the oracle on ALL candidates: regenerate the whole set on this GPU (identical by construction), copy it to the host
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import pyoracle as po

    t0 = time.perf_counter()
    best_c, best_i = float("inf"), -1
    CH = 125_000
    for c0 in range(0, TOTAL, CH):
        p_, t_ = ctx.generate_candidates_batch(CH, K, D, seed=SEED, first_index=c0, pos_min=[-10.0] * D,
                                               pos_max=[10.0] * D, v_max=3.0, a_max=5.0, layout="aos")
        _, cst = po.solve_canonical_batch(p_.cpu().numpy(), t_.cpu().numpy(), n_threads=os.cpu_count() or 1)
        j = int(np.argmin(cst))
        if cst[j] < best_c:
            best_c, best_i = float(cst[j]), c0 + j
    line["oracle"] = {"cost": best_c, "candidate": best_i, "seconds": time.perf_counter() - t0,
                      "threads": os.cpu_count(), "index_equal": best_i == idx,
                      "cost_rel_err": abs(best_c - cost) / best_c}
    line["ok"] = bool(best_i == idx and abs(best_c - cost) <= 1e-9 * best_c)
if rank == 0:
    print(json.dumps(line))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

#!/usr/bin/env python
"""2..8-rank check of the sweep collective (run under torchrun on one box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/nccl_argmin_check.py
Every rank solves its contiguous shard of a 1M-candidate sweep (BASELINE config 5, scaled by
--total), then the global argmin is formed three times: by mtg_argmin_allgather (the library's own
NCCL communicator, bootstrapped with a unique id that rank 0 broadcasts), by mtg_best_allgather
(the same exchange without a host round trip) and by sweep.gather_argmin (torch.distributed). All must agree with each other on every rank and with
the argmin of the concatenated costs gathered on rank 0."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mav_tube_trajectory_generation_b200 as m  # noqa: E402
from mav_tube_trajectory_generation_b200 import sweep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total", type=int, default=1_000_000)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = m.Context(local)
    ids = [ctx.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.nccl_init(ids[0], rank, world)

    start, count = sweep.shard_range(args.total, rank, world)
    pos, times = bench.make_workload(count, seed=1234 + rank)
    sol = ctx.solve_batch(torch.from_numpy(pos).cuda(), torch.from_numpy(times).cuda())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c1 = ctx.argmin_allgather(sol["cost"], status=sol["status"], global_offset=start)       # warm
    ev0.record()
    c1 = ctx.argmin_allgather(sol["cost"], status=sol["status"], global_offset=start)
    ev1.record()
    torch.cuda.synchronize()
    best = ctx.argmin_batch(sol["cost"], status=sol["status"], global_offset=start)
    c2 = sweep.gather_argmin(best=best)
    c3 = ctx.decode_best(ctx.best_allgather(best))       # mtg_best_allgather: gather + fold on the device
    # ground truth: all costs on rank 0
    sizes = [sweep.shard_range(args.total, r, world)[1] for r in range(world)]
    parts = [torch.empty(s, dtype=torch.float64, device="cuda") for s in sizes]
    dist.all_gather(parts, sol["cost"]) if len(set(sizes)) == 1 else None
    ok = c1 == c2 and c1 == c3
    if len(set(sizes)) == 1:
        allc = torch.cat(parts).cpu().numpy()
        want = (float(allc.min()), int(np.flatnonzero(allc == allc.min())[0]))
        ok = ok and c1 == want
    flags = [None] * world
    dist.all_gather_object(flags, (ok, c1))
    if rank == 0:
        print(json.dumps({"world": world, "total": args.total, "all_ranks_agree": all(f[0] for f in flags)
                          and len({f[1] for f in flags}) == 1, "argmin": c1,
                          "allgather_call_ms": ev0.elapsed_time(ev1)}))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

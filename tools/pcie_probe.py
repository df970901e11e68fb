#!/usr/bin/env python
"""Host <-> device copy bandwidth of every rank alone and of all ranks at once (pinned memory, one
process per GPU under torchrun): the platform ceiling of the end-to-end (MTG_MEM_HOST) path.
Run: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py"""
import json
import os
import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 256 << 20
h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
d_out = torch.ones(nbytes, dtype=torch.uint8, device="cuda")
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(up, dn, reps=8):
    """GB/s of `reps` copies of 256 MB in the chosen direction(s), both directions concurrently when both are set."""
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    s_up.wait_event(ev[0])
    s_dn.wait_event(ev[0])
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s_up):
                d_in.copy_(h_in, non_blocking=True)
        if dn:
            with torch.cuda.stream(s_dn):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s_up)
    torch.cuda.current_stream().wait_stream(s_dn)
    ev[1].record()
    torch.cuda.synchronize()
    return reps * nbytes / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9


run(True, True, 2)  # warm-up
res = {}
for name, (up, dn) in {"h2d": (True, False), "d2h": (False, True), "both_each_direction": (True, True)}.items():
    alone = 0.0
    for r in range(world):  # every rank in turn while the others idle
        barrier()
        if r == rank:
            alone = run(up, dn)
    barrier()
    together = run(up, dn)  # all ranks at once
    barrier()
    t = torch.tensor([alone, together], dtype=torch.float64, device="cuda")
    if world > 1:
        g = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(g, t)
    else:
        g = [t]
    res[name] = {"alone_gbs_per_rank": [round(float(x[0]), 1) for x in g],
                 "concurrent_gbs_per_rank": [round(float(x[1]), 1) for x in g],
                 "concurrent_gbs_sum": round(sum(float(x[1]) for x in g), 1)}
if rank == 0:
    print(json.dumps({"n_gpus": world, "bytes_per_copy": nbytes, "unit": "GB/s per direction", **res}))
if world > 1:
    dist.destroy_process_group()

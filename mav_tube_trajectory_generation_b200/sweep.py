"""Host-side plumbing of the sharded candidate sweep (BASELINE config 5).

Every trajectory is independent, so the batch is split into contiguous ranges, one per rank
(one process per GPU), and the only exchange is the final argmin: each rank contributes one
{cost, global index} pair. `gather_argmin` moves the pairs with torch.distributed (NCCL on
the GPUs, gloo in the CPU tests); C++ hosts use mtg_argmin_allgather of the C ABI instead.
Nothing here computes trajectories: the per-rank pair comes from mtg_argmin_batch.
"""
from __future__ import annotations

import struct
from typing import Iterable, Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [start, start + count) of rank `rank`: [g*B/G, (g+1)*B/G)."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request")
    start = (total * rank) // world
    stop = (total * (rank + 1)) // world
    return start, stop - start


def merge_argmin(pairs: Iterable[Tuple[float, int]]) -> Tuple[float, int]:
    """Lowest cost wins, ties go to the lowest global index; idx < 0 (an empty shard or one whose
    solves all failed) and NaN costs never win. Returns (inf, -1) if nothing qualifies."""
    best_c, best_i = float("inf"), -1
    for c, i in pairs:
        if i < 0 or c != c:
            continue
        if best_i < 0 or c < best_c or (c == best_c and i < best_i):
            best_c, best_i = c, i
    return best_c, best_i


def _pack(cost: float, idx: int):
    return [struct.unpack("<q", struct.pack("<d", float(cost)))[0], int(idx)]


def _unpack(bits: int, idx: int) -> Tuple[float, int]:
    return struct.unpack("<d", struct.pack("<q", int(bits)))[0], int(idx)


def gather_argmin(local_cost=None, local_idx=None, best=None, group=None, device=None) -> Tuple[float, int]:
    """All-gather of one 16-byte pair per rank and the final selection; every rank returns the
    same (cost, global index). Pass either the device pair `best` produced by
    Context.argmin_batch (a 2-element int64 CUDA tensor: cost bits, index) or host scalars."""
    import torch
    import torch.distributed as dist

    if best is None:
        best = torch.tensor(_pack(local_cost, local_idx), dtype=torch.int64, device=device or "cpu")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        h = best.cpu().tolist()
        return merge_argmin([_unpack(h[0], h[1])])
    world = dist.get_world_size(group)
    out = torch.empty(2 * world, dtype=torch.int64, device=best.device)
    dist.all_gather_into_tensor(out, best.contiguous(), group=group)
    h = out.cpu().tolist()
    return merge_argmin(_unpack(h[2 * r], h[2 * r + 1]) for r in range(world))

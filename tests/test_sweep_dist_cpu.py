"""CPU tests of the multi-rank plumbing of the sharded sweep (world_size 2, gloo): shard ranges
and the argmin gather. No trajectory is computed here: the per-rank {cost, index} pairs are
inputs (on the GPUs they come from mtg_argmin_batch)."""
import os
import socket

import numpy as np
import pytest

from mav_tube_trajectory_generation_b200 import sweep


def test_shard_ranges_cover_the_batch():
    for total in (0, 1, 7, 65536, 1_000_000):
        for world in (1, 2, 3, 8):
            got = [sweep.shard_range(total, r, world) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == total
            for (s0, c0), (s1, _) in zip(got, got[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in got) - min(c for _, c in got) <= 1
    assert sweep.shard_range(1_000_000, 3, 8) == (375_000, 125_000)   # BASELINE config 5
    with pytest.raises(ValueError):
        sweep.shard_range(10, 2, 2)


def test_merge_rules():
    assert sweep.merge_argmin([(3.0, 5), (2.0, 9), (2.0, 7)]) == (2.0, 7)      # tie -> lower index
    assert sweep.merge_argmin([(float("nan"), 1), (4.0, -1), (5.0, 2)]) == (5.0, 2)
    assert sweep.merge_argmin([]) == (float("inf"), -1)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, costs, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, count = sweep.shard_range(len(costs), rank, world)
        mine = np.asarray(costs[start:start + count])
        ok = ~np.isnan(mine)
        if ok.any():
            j = int(np.flatnonzero(mine == np.nanmin(mine))[0])     # serial scan: first minimum
            pair = (float(mine[j]), start + j)
        else:
            pair = (float("inf"), -1)
        q.put((rank, sweep.gather_argmin(*pair)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["plain", "tie_across_ranks", "empty_rank"])
def test_gather_argmin_world_2(case):
    import torch.multiprocessing as mp

    rng = np.random.RandomState(7)
    costs = rng.uniform(1.0, 2.0, size=1001)
    if case == "tie_across_ranks":
        costs[900] = costs[100] = 0.5            # same cost on both ranks: the lower index wins
    if case == "empty_rank":
        costs[:] = np.nan
        costs[700] = 1.25                        # rank 0 has nothing to offer
    want = (float(np.nanmin(costs)), int(np.flatnonzero(costs == np.nanmin(costs))[0]))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, costs.tolist(), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == want and got[1] == want

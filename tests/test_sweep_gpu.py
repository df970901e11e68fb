"""-m gpu tests of the candidate-sweep reduction: mtg_argmin_batch / mtg_argmin_allgather
(single rank here; the 2-rank NCCL path is exercised by tools/nccl_argmin_check.py under
torchrun, and its host plumbing by tests/test_sweep_dist_cpu.py on gloo)."""
import numpy as np
import pytest

from gpu_util import ctx, dev, host, random_problems, soa

pytestmark = pytest.mark.gpu


def ref_argmin(cost, status=None, offset=0):
    ok = ~np.isnan(cost)
    if status is not None:
        ok &= status == 0
    if not ok.any():
        return float("inf"), -1
    m = cost[ok].min()
    return float(m), offset + int(np.flatnonzero(ok & (cost == m))[0])


@pytest.mark.parametrize("n", [1, 31, 257, 65536, 1_000_003])
def test_argmin_matches_serial_scan(n):
    import torch

    rng = np.random.RandomState(n)
    cost = rng.uniform(1.0, 100.0, size=n)
    if n > 100:
        cost[rng.randint(0, n, size=20)] = np.nan
        lo = cost[~np.isnan(cost)].min() * 0.5
        cost[[n - 3, 17, n // 2]] = lo                    # ties: the lowest index wins
    c = ctx()
    best = c.argmin_batch(dev(cost), global_offset=1000)
    assert c.decode_best(best) == ref_argmin(cost, None, 1000)
    # failed solves (status != 0) never win
    status = np.zeros(n, dtype=np.int32)
    if n > 100:
        status[17] = 2
    best = c.argmin_batch(dev(cost), status=dev(status), global_offset=0)
    assert c.decode_best(best) == ref_argmin(cost, status, 0)
    # the gather entry point without a communicator is the local argmin
    assert c.argmin_allgather(dev(cost), status=dev(status), global_offset=5) == ref_argmin(cost, status, 5)
    torch.cuda.synchronize()


def test_running_argmin_and_empty():
    c = ctx()
    rng = np.random.RandomState(3)
    chunks = [rng.uniform(0, 10, size=5000) for _ in range(6)]
    best = None
    for k, ch in enumerate(chunks):
        best = c.argmin_batch(dev(ch), global_offset=5000 * k, best=best, accumulate=True)
    assert c.decode_best(best) == ref_argmin(np.concatenate(chunks))
    nothing = np.full(100, np.nan)
    assert c.decode_best(c.argmin_batch(dev(nothing))) == (float("inf"), -1)
    assert c.argmin_allgather(dev(nothing)) == (float("inf"), -1)


def test_sweep_argmin_equals_oracle(po):
    """BASELINE config 5 at a size the oracle finishes in seconds: the index and cost of the best
    candidate equal the oracle's (cost parity 1e-9; the winner is well separated)."""
    B = 20000
    pos, times = random_problems(po, 200, 10, 3, seed0=123)
    rng = np.random.RandomState(5)
    pos = np.repeat(pos, B // 200, axis=0) + rng.normal(0, 1.0, size=(B, 11, 3))
    times = np.repeat(times, B // 200, axis=0) * rng.uniform(0.9, 1.3, size=(B, 10))
    c = ctx()
    sol = c.solve_batch(dev(soa(pos)), dev(soa(times)))
    cost_gpu = host(sol["cost"])
    got_c, got_i = c.argmin_allgather(sol["cost"], status=sol["status"])
    assert (got_c, got_i) == ref_argmin(cost_gpu, host(sol["status"]))
    _, ref_cost = po.solve_canonical_batch(pos, times, n_threads=8)
    order = np.argsort(ref_cost)
    assert got_i == order[0]
    assert abs(got_c - ref_cost[order[0]]) <= 1e-9 * ref_cost[order[0]]
    assert ref_cost[order[1]] - ref_cost[order[0]] > 1e-6 * ref_cost[order[0]]   # a real winner


def test_full_size_sweep_strided_sample_vs_oracle(po):
    """BASELINE configs[3] at FULL size — the 1,000,000-trajectory x ~1000-sample evaluateRange + v/a/tube
    sweep that bench.py times — with every 4096th trajectory compared against the oracle: sample count,
    sampling times and segment index bit-exact (the reference's serial recurrence, TRAJ_C:74-134), values
    to 1e-12 of the polynomial's scale, flags identical except within 1e-9 of a limit. The same check runs
    inside the bench line (sweep.parity_sample)."""
    import torch

    import bench

    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~35 GB of free HBM")
    c = ctx()
    out = bench.sweep_section(c, 6454.6, batch=1_000_000, time_it=False)
    ps = out["parity_sample"]
    assert ps["trajectories"] == 245 and ps["ok"], ps
    assert ps["count_mismatch"] == 0 and ps["time_mismatch"] == 0 and ps["segment_mismatch"] == 0
    assert ps["value_err_max_rel"] <= 1e-12 and ps["flag_mismatch_away_from_limits"] == 0
    assert ps["subset_rows_bit_identical_to_full_run"]


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_device_candidate_generator_matches_host_twin(po, layout):
    """mtg_generate_candidates_batch (Philox4x32-10 keyed by the candidate index; createRandomVertices' 0.2 m rejection
    rule VTX_C:65-72; Nfabian times VTX_C:252-269) against the numpy twin: positions bit for bit, times to 2 ulp
    (device exp vs libm exp); independent of the batch split (first_index) and of the layout."""
    from gpu_util import aos

    c = ctx()
    B, K, D, seed = 5000, 10, 3, 0xB200
    p, t = c.generate_candidates_batch(B, K, D, seed, layout=layout)
    conv = aos if layout == "soa" else (lambda x: x)
    pos, times = conv(host(p)), conv(host(t))
    want_p, want_t = po.generate_candidates(B, K, D, seed)
    assert np.array_equal(pos, want_p)
    assert np.abs(times - want_t).max() <= 4e-16 * np.abs(want_t).max()
    assert np.all(np.linalg.norm(np.diff(pos, axis=1), axis=2) > 0.2) and pos.min() >= -10 and pos.max() < 10
    # a shard that starts at candidate 3000 reproduces the tail of the full batch
    p2, t2 = c.generate_candidates_batch(2000, K, D, seed, first_index=3000, layout=layout)
    assert np.array_equal(conv(host(p2)), pos[3000:]) and np.array_equal(conv(host(t2)), times[3000:])
    # a tight box makes the rejection rule fire (|pos - last| <= 0.2 has probability ~3 % per draw in a 1 m box)
    p3, _ = c.generate_candidates_batch(4000, 4, 3, 7, pos_min=0.0, pos_max=1.0, layout=layout)
    w3, _ = po.generate_candidates(4000, 4, 3, 7, pos_min=0.0, pos_max=1.0)
    assert np.array_equal(conv(host(p3)), w3)
    assert np.all(np.linalg.norm(np.diff(w3, axis=1), axis=2) > 0.2)
    # 2-D and 1-D candidates use one Philox block per draw
    for d_ in (1, 2):
        p4, _ = c.generate_candidates_batch(300, 3, d_, 11, layout=layout)
        w4, _ = po.generate_candidates(300, 3, d_, 11)
        assert np.array_equal(conv(host(p4)), w4)

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


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_fused_solve_argmin_equals_the_two_launches(po, layout):
    """mtg_solve_argmin_batch = mtg_solve_batch + mtg_argmin_batch, bit for bit: same outputs, same winner under the
    same total order (ties -> lower global index, failed solves and NaN costs never win), running best over several
    batches, outputs optional, and the K = 1 shape that takes the two-launch route inside the library."""
    import torch

    c = ctx()
    B, K = 5000, 10                                   # not a multiple of the CTA's 64 trajectories
    pos, times = c.generate_candidates_batch(B, K, 3, seed=5, layout=layout)
    def col(x, b):                                    # view of candidate b
        return x[b] if layout == "aos" else x[..., b]
    col(times, 17)[0] = -1.0                          # a failed solve ...
    plain = c.solve_batch(pos, times, layout=layout)
    cost0 = host(plain["cost"]).copy()
    st0 = host(plain["status"])
    assert st0[17] != 0
    best_b = int(np.flatnonzero((st0 == 0) & (cost0 == cost0[st0 == 0].min()))[0])
    # ... and an exact tie with the best candidate at a HIGHER index: the lower one must win
    twin = B - 3 if best_b != B - 3 else B - 4
    col(pos, twin).copy_(col(pos, best_b))
    col(times, twin).copy_(col(times, best_b))
    plain = c.solve_batch(pos, times, layout=layout)
    want = ref_argmin(host(plain["cost"]), host(plain["status"]), offset=1000)
    assert want[1] == 1000 + min(best_b, twin)
    out = {"coeffs": torch.empty_like(plain["coeffs"]), "cost": torch.empty_like(plain["cost"]),
           "status": torch.empty_like(plain["status"])}
    best = c.solve_argmin_batch(pos, times, layout=layout, global_offset=1000, out=out)
    assert c.decode_best(best) == want
    for k in out:
        assert torch.equal(out[k], plain[k]), k
    # no outputs at all, accumulate over three batches with their own offsets (the middle one holds the winner)
    run = torch.zeros(2, dtype=torch.int64, device="cuda")
    pos2, times2 = c.generate_candidates_batch(B, K, 3, seed=6, layout=layout)
    p2 = c.solve_batch(pos2, times2, layout=layout)
    for n, (p_, t_, off) in enumerate(((pos2, times2, 0), (pos, times, 7000), (pos2, times2, 20000))):
        c.solve_argmin_batch(p_, t_, layout=layout, global_offset=off, best=run, accumulate=n > 0)
    a = ref_argmin(host(p2["cost"]), host(p2["status"]), 0)
    b = ref_argmin(host(plain["cost"]), host(plain["status"]), 7000)
    assert c.decode_best(run) == min(a, b)
    # the fused argmin agrees with mtg_argmin_batch itself
    assert c.decode_best(c.argmin_batch(plain["cost"], plain["status"], global_offset=1000)) == want
    # K = 1: two launches inside the library, scratch for the cost the caller did not ask for
    pos1, times1 = c.generate_candidates_batch(300, 1, 3, seed=7, layout=layout)
    p1 = c.solve_batch(pos1, times1, layout=layout)
    assert c.decode_best(c.solve_argmin_batch(pos1, times1, layout=layout)) == \
        ref_argmin(host(p1["cost"]), host(p1["status"]))


def test_overlapped_solve_train_is_bit_identical():
    """mtg_set_solve_overlap: a train of solves launched with programmatic stream serialization (each may start
    while the previous one drains, and waits for it before its first store) writes the SAME outputs as ordinary
    launches — also when every call reuses one output buffer — and the fused running argmin is the same."""
    import torch

    c = ctx()
    B, K, n = 20000, 10, 6
    batches = [c.generate_candidates_batch(B, K, 3, seed=40 + i) for i in range(n)]
    torch.cuda.synchronize()
    want, want_best = [], torch.zeros(2, dtype=torch.int64, device="cuda")
    for i, (p, t) in enumerate(batches):
        want.append(c.solve_batch(p, t))
        c.argmin_batch(want[-1]["cost"], status=want[-1]["status"], global_offset=i * B, best=want_best,
                       accumulate=i > 0)
    torch.cuda.synchronize()
    try:
        c.set_solve_overlap(True)
        for rep in range(3):
            # one output buffer for the whole train: only the last solve's results may be in it afterwards
            out = {k: torch.empty_like(v) for k, v in want[0].items() if v is not None}
            best = torch.zeros(2, dtype=torch.int64, device="cuda")
            for i, (p, t) in enumerate(batches):
                c.solve_argmin_batch(p, t, out=out, global_offset=i * B, best=best, accumulate=i > 0)
            torch.cuda.synchronize()
            for k in out:
                assert torch.equal(out[k], want[-1][k]), k
            assert torch.equal(best, want_best)
            # separate outputs, plain mtg_solve_batch
            outs = [c.solve_batch(p, t) for p, t in batches]
            torch.cuda.synchronize()
            for a, b in zip(outs, want):
                assert torch.equal(a["coeffs"], b["coeffs"]) and torch.equal(a["cost"], b["cost"])
    finally:
        c.set_solve_overlap(False)


def test_fused_argmin_corner_cases():
    """Nothing qualifies (every solve fails): the pair stays {inf, -1}; a later batch still wins; the K = 1 shape
    with the overlap option on takes the two-launch route and gives the same pair."""
    import torch

    c = ctx()
    B, K = 1000, 4
    pos, times = c.generate_candidates_batch(B, K, 3, seed=90)
    bad = -times.abs()
    st = torch.empty((B,), dtype=torch.int32, device="cuda")
    best = c.solve_argmin_batch(pos, bad, out={"status": st})
    cost, idx = c.decode_best(best)
    assert idx == -1 and cost == float("inf") and int((st != 0).sum()) == B
    c.solve_argmin_batch(pos, times, global_offset=5000, best=best, accumulate=True)
    plain = c.solve_batch(pos, times)
    assert c.decode_best(best) == ref_argmin(host(plain["cost"]), host(plain["status"]), 5000)
    pos1, times1 = c.generate_candidates_batch(500, 1, 3, seed=91)
    p1 = c.solve_batch(pos1, times1)
    try:
        c.set_solve_overlap(True)
        got = c.decode_best(c.solve_argmin_batch(pos1, times1))
        again = c.solve_batch(pos1, times1)
    finally:
        c.set_solve_overlap(False)
    assert got == ref_argmin(host(p1["cost"]), host(p1["status"]))
    assert torch.equal(again["coeffs"], p1["coeffs"]) and torch.equal(again["cost"], p1["cost"])


def test_overlapped_sweep_with_device_generator():
    """The sweep loop of INTEGRATION.md with mtg_set_solve_overlap on: batch it + 1 is generated (into the other
    buffer pair) BEFORE the solve of batch it is launched, so no solve reads what the operation right before it
    wrote. Same running best as the plain loop."""
    import torch

    c = ctx()
    B, K, n = 30000, 10, 8

    def sweep(overlap):
        best = torch.zeros(2, dtype=torch.int64, device="cuda")
        bufs = [c.generate_candidates_batch(B, K, 3, seed=123, first_index=0), None]
        c.set_solve_overlap(overlap)
        try:
            for it in range(n):
                if it + 1 < n:
                    if bufs[(it + 1) % 2] is None:
                        bufs[(it + 1) % 2] = tuple(torch.empty_like(x) for x in bufs[0])
                    nxt = c.generate_candidates_batch(B, K, 3, seed=123, first_index=(it + 1) * B)
                    for dst, src in zip(bufs[(it + 1) % 2], nxt):   # generator output lands in the other pair
                        dst.copy_(src)
                p, t = bufs[it % 2]
                c.solve_argmin_batch(p, t, global_offset=it * B, best=best, accumulate=it > 0)
            torch.cuda.synchronize()
        finally:
            c.set_solve_overlap(False)
        return c.decode_best(best)

    assert sweep(True) == sweep(False)


def test_fused_argmin_at_a_million_candidates():
    """Full size of a sweep step (BASELINE configs[4]): 1,000,000 candidates, 15,625 CTAs folding into one pair."""
    c = ctx()
    B = 1_000_000
    pos, times = c.generate_candidates_batch(B, 10, 3, seed=31337)
    plain = c.solve_batch(pos, times, want_coeffs=False)
    want = c.decode_best(c.argmin_batch(plain["cost"], plain["status"], global_offset=10**12))
    assert c.decode_best(c.solve_argmin_batch(pos, times, global_offset=10**12)) == want
    try:
        c.set_solve_overlap(True)
        assert c.decode_best(c.solve_argmin_batch(pos, times, global_offset=10**12)) == want
    finally:
        c.set_solve_overlap(False)

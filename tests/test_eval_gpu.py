"""-m gpu parity tests of the sampled path: mtg_eval_range_batch (E4), mtg_eval_at_batch
(E3), mtg_max_time_batch and mtg_feasibility_batch (E5 sampled, T1) against the oracle.

Bars (BASELINE.json north_star): sample count, sampling times and segment lookup are
BIT-EXACT (they come from the reference's serial fp64 recurrence); sample values are
fp64 Horner sums compared at 1e-12 of the polynomial's magnitude scale (the reference
itself is FMA-contraction dependent); feasibility flags identical except for samples
within 1e-9 (relative) of a limit.
"""
import numpy as np
import pytest

from gpu_util import aos, ctx, dev, host, random_problems, soa

pytestmark = pytest.mark.gpu
N = 10


@pytest.fixture(scope="module")
def solved(po):
    """64 random config-C trajectories solved by the oracle: coeffs [B,K,D,N], times [B,K], pos."""
    pos, times = random_problems(po, 64, 10, 3, seed0=400)
    coeffs, _ = po.solve_canonical_batch(pos, times, n_threads=8)
    return pos, times, coeffs


def max_times(times):
    out = np.zeros(len(times))
    for b, t in enumerate(times):
        acc = 0.0
        for x in t:
            acc += x
        out[b] = acc
    return out


def value_tol(coeffs, times, derivative):
    """1e-12 x sum_j |B(d,j) c_j| T^(j-d): the magnitude scale of a Horner evaluation."""
    B, K, D, n = coeffs.shape
    j = np.arange(n)
    fall = np.ones(n)
    for q in range(derivative):
        fall = fall * np.maximum(j - q, 0)
    expo = np.maximum(j - derivative, 0)
    scale = (np.abs(coeffs) * fall * times[:, :, None, None] ** expo).sum(axis=-1)   # [B,K,D]
    return 1e-12 * np.maximum(scale.max(axis=(1, 2)), 1e-300)


def run_range(coeffs, times, t0, t1, dt, derivative, layout, device=True, S=1100):
    c = ctx()
    if layout == "soa":
        cc, tt = soa(coeffs), soa(times)
    else:
        cc, tt = np.ascontiguousarray(coeffs), np.ascontiguousarray(times)
    if device:
        cc, tt = dev(cc), dev(tt)
    r = c.eval_range_batch(cc, tt, t0, t1, dt, derivative=derivative, max_samples=S, layout=layout,
                           want_times=True, want_segments=True)
    if device:
        import torch

        torch.cuda.synchronize()
    conv = aos if layout == "soa" else (lambda x: x)
    return (conv(host(r["samples"])), conv(host(r["sampling_times"])), conv(host(r["segment_idx"])),
            host(r["n_samples"]), host(r["status"]))


@pytest.mark.parametrize("layout", ["soa", "aos"])
@pytest.mark.parametrize("derivative", [0, 1, 2, 4])
def test_eval_range_vs_oracle(po, solved, layout, derivative):
    pos, times, coeffs = solved
    tmax = max_times(times)
    dt = tmax / 1000
    samples, st, seg, n, status = run_range(coeffs, times, 0.0, tmax, dt, derivative, layout)
    tol = value_tol(coeffs, times, derivative)
    assert np.all(status == 0)
    counts = set()
    for b in range(len(pos)):
        ref = po.traj_evaluate_range(coeffs[b], times[b], 0.0, tmax[b], dt[b], derivative)
        k = ref[0].shape[0]
        counts.add(k)
        assert n[b] == k
        assert np.array_equal(st[b, :k], ref[1])          # accumulated_time: bit-exact
        assert np.array_equal(seg[b, :k], ref[2])         # segment lookup: bit-exact
        assert np.abs(samples[b, :k] - ref[0]).max() <= tol[b]
    assert counts <= {1000, 1001} and len(counts) == 2    # both counts occur


def test_max_time(po, solved):
    _, times, _ = solved
    got = host(ctx().max_time_batch(dev(soa(times))))
    assert np.array_equal(got, max_times(times))
    got = ctx().max_time_batch(np.ascontiguousarray(times), layout="aos")
    assert np.array_equal(got, max_times(times))


def test_eval_range_grid_aligned_indexing(po):
    """Segment times on a 0.1 s grid with dt = 0.01: the case where k*dt + cumsum lookup
    disagrees with the reference's recurrence for ~0.4 % of the samples."""
    rng = np.random.RandomState(3)
    B, K, D = 48, 10, 3
    pos = rng.uniform(-10, 10, size=(B, K + 1, D))
    times = np.round(rng.uniform(0.5, 4.0, size=(B, K)), 1)
    coeffs, _ = po.solve_canonical_batch(pos, times, n_threads=8)
    tmax = max_times(times)
    S = int(tmax.max() / 0.01) + 8
    samples, st, seg, n, status = run_range(coeffs, times, 0.0, tmax, 0.01, 0, "soa", S=S)
    naive_mismatch = 0
    for b in range(B):
        ref = po.traj_evaluate_range(coeffs[b], times[b], 0.0, tmax[b], 0.01, 0)
        k = ref[0].shape[0]
        assert n[b] == k and status[b] == 0
        assert np.array_equal(st[b, :k], ref[1]) and np.array_equal(seg[b, :k], ref[2])
        naive = np.searchsorted(np.cumsum(times[b]), np.arange(k) * 0.01, side="right")
        naive_mismatch += int((np.minimum(naive, K - 1) != ref[2]).sum())
    assert naive_mismatch > 0   # the test really exercises the recurrence


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_eval_range_windows_and_errors(po, solved, layout):
    pos, times, coeffs = solved
    B = len(pos)
    tmax = max_times(times)
    # a window that starts inside a segment: the reference restarts its loop counter at that
    # segment's start (TRAJ_C:110-114), shifting the window rather than clipping it
    t0 = 0.37 * tmax
    t1 = 0.61 * tmax
    dt = tmax / 777
    samples, st, seg, n, status = run_range(coeffs, times, t0, t1, dt, 1, layout)
    tol = value_tol(coeffs, times, 1)
    for b in range(B):
        ref = po.traj_evaluate_range(coeffs[b], times[b], t0[b], t1[b], dt[b], 1)
        k = ref[0].shape[0]
        assert n[b] == k and status[b] == 0
        assert np.array_equal(st[b, :k], ref[1]) and np.array_equal(seg[b, :k], ref[2])
        assert np.abs(samples[b, :k] - ref[0]).max() <= tol[b]
    # errors: start beyond the end, start exactly at the end (UB in the reference), dt <= 0
    t0 = np.zeros(B)
    t0[1] = tmax[1] + 1.0
    t0[2] = tmax[2]
    dtv = (tmax / 100).copy()
    dtv[3] = 0.0
    _, _, _, n, status = run_range(coeffs, times, t0, tmax, dtv, 0, layout)
    assert status[1] == 4 and status[2] == 4 and status[3] == 4
    assert n[1] == 0 and n[2] == 0 and n[3] == 0
    assert np.all(status[4:] == 0) and status[0] == 0
    assert po.traj_evaluate_range(coeffs[1], times[1], t0[1], tmax[1], dtv[1], 0) is None
    # empty window
    _, _, _, n, status = run_range(coeffs, times, 0.0, 0.0, 0.01, 0, layout)
    assert np.all(n == 0) and np.all(status == 0)
    # truncation
    samples, st, seg, n, status = run_range(coeffs, times, 0.0, tmax, tmax / 1000, 0, layout, S=100)
    assert np.all(n == 100) and np.all(status == 8)


def test_eval_range_host_mode_and_other_shapes(po):
    import os

    for K, D in ((1, 1), (3, 2), (7, 4)):
        pos, times = random_problems(po, 50, K, D, seed0=900 + K)
        coeffs, _ = po.solve_canonical_batch(pos, times, n_threads=8)
        tmax = max_times(times)
        a = run_range(coeffs, times, 0.0, tmax, tmax / 200, 2, "soa", S=210)
        os.environ["MTG_HOST_CHUNK"] = "16"
        try:
            b = run_range(coeffs, times, 0.0, tmax, tmax / 200, 2, "soa", device=False, S=210)
            c = run_range(coeffs, times, 0.0, tmax, tmax / 200, 2, "aos", device=False, S=210)
        finally:
            del os.environ["MTG_HOST_CHUNK"]
        for bidx in range(50):
            k = a[3][bidx]
            ref = po.traj_evaluate_range(coeffs[bidx], times[bidx], 0.0, tmax[bidx], tmax[bidx] / 200, 2)
            assert k == ref[0].shape[0] == b[3][bidx] == c[3][bidx]
            for x in (a, b, c):
                assert np.array_equal(x[1][bidx, :k], ref[1]) and np.array_equal(x[2][bidx, :k], ref[2])
            assert np.array_equal(a[0][bidx, :k], b[0][bidx, :k])
            assert np.array_equal(a[0][bidx, :k], c[0][bidx, :k])
            assert np.abs(a[0][bidx, :k] - ref[0]).max() <= value_tol(coeffs, times, 2)[bidx]


def test_eval_range_generic_n(po):
    """N = 8 polynomials go through the zero-padded NT = 12 kernel."""
    rng = np.random.RandomState(8)
    B, K, D, n8 = 20, 4, 3, 8
    coeffs = rng.uniform(-1, 1, size=(B, K, D, n8))
    times = rng.uniform(0.5, 3.0, size=(B, K))
    tmax = max_times(times)
    c = ctx()
    r = c.eval_range_batch(dev(soa(coeffs)), dev(soa(times)), 0.0, tmax, tmax / 100, derivative=1,
                           max_samples=110, want_times=True, want_segments=True)
    samples, n = aos(host(r["samples"])), host(r["n_samples"])
    for b in range(B):
        ref = po.traj_evaluate_range(coeffs[b], times[b], 0.0, tmax[b], tmax[b] / 100, 1)
        assert n[b] == ref[0].shape[0]
        assert np.abs(samples[b, :n[b]] - ref[0]).max() <= 1e-12 * 8 * 7 * 3.0 ** 7


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_eval_at_vs_oracle(po, solved, layout):
    pos, times, coeffs = solved
    B, K = times.shape
    tmax = max_times(times)
    rng = np.random.RandomState(2)
    M = 40
    t = rng.uniform(0, 1, size=(B, M)) * tmax[:, None]
    t[:, 0] = 0.0
    t[:, 1] = tmax                      # exactly the end: last segment (TRAJ_C:63-67)
    t[:, 2] = tmax + 0.5                # beyond: zeros + error
    acc = np.zeros(B)
    for i in range(3):                  # exactly on vertices: right-hand segment
        acc = acc + times[:, i]
        t[:, 3 + i] = acc
    c = ctx()
    for derivative in (0, 1, 3):
        if layout == "soa":
            r = c.eval_at_batch(dev(soa(coeffs)), dev(soa(times)), dev(soa(t)), derivative=derivative)
            out, seg = aos(host(r["out"])), aos(host(r["segment_idx"]))
        else:
            r = c.eval_at_batch(coeffs, times, np.ascontiguousarray(t), derivative=derivative, layout="aos")
            out, seg = r["out"], r["segment_idx"]
        status = host(r["status"])
        tol = value_tol(coeffs, times, derivative)
        for b in range(B):
            for m in range(M):
                want, s = po.traj_evaluate(coeffs[b], times[b], t[b, m], derivative)
                assert seg[b, m] == s
                assert np.abs(out[b, m] - want).max() <= tol[b]
            assert seg[b, 2] == -1 and np.all(out[b, 2] == 0.0) and status[b] == 4
            assert seg[b, 1] == K - 1 and seg[b, 3] == 1 and seg[b, 4] == 2


def run_feas(coeffs, times, pos, radii, t0, t1, dt, layout, device=True, S=1100, v_max=3.0, a_max=5.0):
    c = ctx()
    conv_in = soa if layout == "soa" else np.ascontiguousarray
    args = [conv_in(coeffs), conv_in(times)]
    p = conv_in(pos) if radii is not None else None
    r_ = conv_in(radii) if radii is not None else None
    if device:
        args = [dev(a) for a in args]
        p = dev(p) if p is not None else None
        r_ = dev(r_) if r_ is not None else None
    r = c.feasibility_batch(args[0], args[1], t0, t1, dt, v_max, a_max, positions=p, radii=r_,
                            max_samples=S, layout=layout, want_samples=True)
    conv = aos if layout == "soa" else (lambda x: x)
    return {k: (conv(host(v)) if k in ("samples", "flags") else host(v)) for k, v in r.items()}


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_feasibility_vs_oracle(po, solved, layout):
    pos, times, coeffs = solved
    B, K = times.shape
    tmax = max_times(times)
    dt = tmax / 1000
    radii = np.full((B, K, 2), 0.15)
    radii[:, :, 0] += np.random.RandomState(1).uniform(0, 2.0, size=(B, K))   # some samples inside
    r = run_feas(coeffs, times, pos, radii, 0.0, tmax, dt, layout)
    rs = run_range(coeffs, times, 0.0, tmax, dt, 0, layout)
    n_in_tube = 0
    for b in range(B):
        ref = po.feasibility_sweep(coeffs[b], times[b], pos[b], radii[b], 3.0, 5.0, 0.0, tmax[b], dt[b])
        k = ref[0].shape[0]
        assert r["n_samples"][b] == k and r["status"][b] == 0
        assert np.array_equal(r["samples"][b, :k], rs[0][b, :k])     # same FMA chain as evaluateRange
        assert np.abs(r["samples"][b, :k] - ref[0]).max() <= value_tol(coeffs, times, 0)[b]
        assert abs(r["max_v"][b] - ref[2]) <= 1e-12 * ref[2] and abs(r["max_a"][b] - ref[3]) <= 1e-12 * ref[3]
        # flags identical except for samples within 1e-9 of a limit
        v, _, _ = po.traj_evaluate_range(coeffs[b], times[b], 0.0, tmax[b], dt[b], 1)
        a, _, _ = po.traj_evaluate_range(coeffs[b], times[b], 0.0, tmax[b], dt[b], 2)
        nv, na = np.sqrt((v ** 2).sum(1)), np.sqrt((a ** 2).sum(1))
        geom = po.tube_geometry(pos[b], radii[b])
        _, _, segs = po.traj_evaluate_range(coeffs[b], times[b], 0.0, tmax[b], dt[b], 0)
        g = geom[segs]
        x = ref[0]
        A = g[:, :9].reshape(-1, 3, 3)
        y = np.einsum("nij,nj->ni", A, x) + g[:, 9:12]
        q = (y ** 2).sum(1)
        along_s = -np.einsum("ni,ni->n", g[:, 12:15], x - g[:, 15:18])
        along_e = np.einsum("ni,ni->n", g[:, 12:15], x - g[:, 18:21])
        near = (np.abs(nv - 3.0) < 3e-9) | (np.abs(na - 5.0) < 5e-9) | \
               (np.abs(q - g[:, 21] ** 2) < 1e-9 * np.maximum(1.0, q)) | (np.abs(along_s) < 1e-9) | \
               (np.abs(along_e) < 1e-9)
        same = r["flags"][b, :k] == ref[1]
        assert np.all(same | near), (b, np.nonzero(~(same | near))[0][:5])
        n_in_tube += int(((ref[1] & 4) != 0).sum())
        want_feasible = bool(np.all(ref[1] == 7))
        if not near.any():
            assert bool(r["feasible"][b]) == want_feasible
    assert n_in_tube > 0
    # without a tube, bit2 is always set and feasibility only depends on v / a
    r2 = run_feas(coeffs, times, pos, None, 0.0, tmax, dt, layout)
    for b in range(B):
        k = r2["n_samples"][b]
        assert np.all(r2["flags"][b, :k] & 4)
        assert np.array_equal(r2["flags"][b, :k] & 3, r["flags"][b, :k] & 3)
        assert bool(r2["feasible"][b]) == bool(np.all(r2["flags"][b, :k] == 7))


def test_feasibility_limits_scale_with_times(po):
    """Slowing a trajectory down by s divides |v| by s and |a| by s^2 (sizes of BASELINE config 4
    are covered by bench_sweep; this is the size-independent property)."""
    pos, times = random_problems(po, 2048, 10, 3, seed0=77)
    c = ctx()
    p, t = dev(soa(pos)), dev(soa(times))
    out = {}
    for s in (1.0, 2.0):
        sol = c.solve_batch(p, t * s)
        tm = c.max_time_batch(t * s)
        r = c.feasibility_batch(sol["coeffs"], t * s, 0.0, tm, tm / 500, 3.0, 5.0, max_samples=510)
        out[s] = (host(r["max_v"]), host(r["max_a"]), host(r["n_samples"]), host(r["status"]))
    assert np.all(out[1.0][3] == 0) and np.all(out[2.0][3] == 0)
    assert np.all((out[1.0][2] == 500) | (out[1.0][2] == 501))
    assert np.allclose(out[2.0][0], out[1.0][0] / 2, rtol=1e-9)
    assert np.allclose(out[2.0][1], out[1.0][1] / 4, rtol=1e-9)


def test_sweep_at_scale_two_kernels_agree(po):
    """BASELINE config 4 at 1/8 size (131,072 trajectories x ~1000 samples): the two independent
    implementations of the sweep — one thread per trajectory with SoA outputs (eval.cuh) and the
    warp-cooperative kernel with trajectory-contiguous outputs (eval_tm.cuh) — replay the same serial
    recurrence and the same FMA chains, so sample counts, flags and every sample agree BIT FOR BIT;
    plus the size-independent facts: 1000 or 1001 samples, the first sample is the start vertex."""
    import torch

    B = 131072
    base_pos, base_t = random_problems(po, 256, 10, 3, seed0=4000)
    rng = np.random.RandomState(11)
    pos = np.repeat(base_pos, B // 256, axis=0) + rng.normal(0, 0.3, size=(B, 11, 3))
    times = np.repeat(base_t, B // 256, axis=0) * rng.uniform(0.7, 1.6, size=(B, 10))
    c = ctx()
    p_soa, t_soa = dev(soa(pos)), dev(soa(times))
    sol = c.solve_batch(p_soa, t_soa)
    coeffs_aos = sol["coeffs"].permute(3, 0, 1, 2).contiguous()
    t_aos = t_soa.t().contiguous()
    p_aos = p_soa.permute(2, 0, 1).contiguous()
    tm = c.max_time_batch(t_soa)
    S = 1010
    radii_aos = torch.full((B, 10, 2), 0.4, dtype=torch.float64, device="cuda")
    radii_soa = radii_aos.permute(1, 2, 0).contiguous()
    a = c.feasibility_batch(coeffs_aos, t_aos, 0.0, tm, tm / 1000, 3.0, 5.0, positions=p_aos, radii=radii_aos,
                            max_samples=S, layout="aos", want_samples=True)
    s = c.feasibility_batch(sol["coeffs"], t_soa, 0.0, tm, tm / 1000, 3.0, 5.0, positions=p_soa, radii=radii_soa,
                            max_samples=S, layout="soa", want_samples=True)
    n = a["n_samples"]
    assert torch.equal(n, s["n_samples"]) and int(a["status"].max()) == 0 and int(s["status"].max()) == 0
    assert bool(((n == 1000) | (n == 1001)).all()) and int((n == 1001).sum()) > 0
    valid = torch.arange(S, device="cuda")[None, :] < n[:, None]                       # [B, S]
    sa, ss = a["samples"], s["samples"].permute(2, 0, 1)                                # [B, S, 3]
    assert torch.equal(torch.where(valid[:, :, None], sa, torch.zeros_like(sa)),
                       torch.where(valid[:, :, None], ss, torch.zeros_like(ss)))
    fa, fs = a["flags"], s["flags"].t()
    assert torch.equal(torch.where(valid, fa, torch.zeros_like(fa)), torch.where(valid, fs, torch.zeros_like(fs)))
    assert torch.equal(a["max_v"], s["max_v"]) and torch.equal(a["max_a"], s["max_a"])
    assert torch.equal(a["feasible"], s["feasible"])
    assert float((sa[:, 0, :] - p_aos[:, 0, :]).abs().max()) < 1e-9                    # starts at vertex 0
    assert int((fa[valid] & 4).sum()) > 0                                               # some samples inside the tube

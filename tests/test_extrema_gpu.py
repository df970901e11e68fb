"""-m gpu parity tests of mtg_extrema_batch (E6 / R1: analytic extrema of the derivative
magnitude) against the oracle, whose root finder IS the reference's Jenkins-Traub (rpoly_ak1.cpp
compiled verbatim into oracle/_ref) behind a restatement of segment.cpp:82-184 /
trajectory.cpp:184-220.

Bars (SURVEY.md section 7.4 "compare values, not root lists"):
  extremum VALUES  : 1e-9 relative to the trajectory's maximum (second-order insensitive to the
                     root, so in practice ~1e-14);
  segment of the max: identical;  time of the max: 1e-6 * T_segment (two different root finders);
  the minimum's time/segment only where the minimum is not a (near-)tie: at rest-to-rest ends
  the magnitude is ~1e-15 at both t = 0 of segment 0 and t = T of the last segment.
"""
import numpy as np
import pytest

from conftest import REFERENCE_PARAMS, make_reference_problem
from gpu_util import aos, ctx, dev, host, random_problems, soa

pytestmark = pytest.mark.gpu
N = 10


def gpu_extrema(coeffs, times, derivative, layout="soa", device=True):
    c = ctx()
    conv_in = soa if layout == "soa" else np.ascontiguousarray
    cc, tt = conv_in(coeffs), conv_in(times)
    if device:
        cc, tt = dev(cc), dev(tt)
    r = c.extrema_batch(cc, tt, derivative, layout=layout, want_segments=True)
    if device:
        import torch

        torch.cuda.synchronize()
    conv = aos if layout == "soa" else (lambda x: x)
    return {k: (conv(host(v)) if k.startswith("seg_") else host(v)) for k, v in r.items()}


def check_against_oracle(po, coeffs, times, der, r, idx):
    for b in idx:
        (mt, mv, ms), (Mt, Mv, Ms) = po.traj_min_max_magnitude(coeffs[b], times[b], der)
        assert r["status"][b] == 0
        assert abs(r["max_value"][b] - Mv) <= 1e-9 * Mv, (b, r["max_value"][b], Mv)
        assert r["max_seg"][b] == Ms
        assert abs(r["max_time"][b] - Mt) <= 1e-6 * times[b][Ms], (b, r["max_time"][b], Mt)
        assert abs(r["min_value"][b] - mv) <= 1e-9 * Mv
        if mv > 1e-6 * Mv:   # a genuine interior minimum, not the ~0 of a rest-to-rest end
            same_place = r["min_seg"][b] == ms and abs(r["min_time"][b] - mt) <= 1e-6 * times[b][ms]
            if not same_place:
                # ... or an equally good one: next to a rest vertex the magnitude is flat to rounding over ~1e-3 T
                # (five vanishing derivatives), and which point of the plateau a root finder reports is its own
                # business — the oracle's magnitude at the reported time must equal its minimum
                t_abs = float(np.sum(times[b][:r["min_seg"][b]]) + r["min_time"][b])
                val = np.sqrt(np.sum(po.traj_evaluate(coeffs[b], times[b], t_abs, der)[0] ** 2))
                assert abs(val - mv) <= 1e-12 * Mv, (b, r["min_time"][b], mt, val, mv)
        # per-segment maxima = what computeMaximumOfMagnitude (LIN_I:455-487) consumes
        t, v, s = po.opt_max_magnitude(coeffs[b], times[b], der)
        assert abs(r["seg_max_value"][b].max() - v) <= 1e-9 * v and int(np.argmax(r["seg_max_value"][b])) == s


@pytest.mark.parametrize("name", list(REFERENCE_PARAMS))
@pytest.mark.parametrize("der", [0, 1, 2, 3])
def test_reference_parameter_sets(po, name, der):
    prob = make_reference_problem(name)
    s = po.solve(N, prob["derivative"], prob["times"], prob["mask"], prob["values"])
    coeffs, times = s.coeffs[None], prob["times"][None]
    r = gpu_extrema(coeffs, times, der)
    check_against_oracle(po, coeffs, times, der, r, [0])
    if der in (1, 2):
        # ExtremaOfMagnitude, TEST_OPT:307-406: the analytic maximum equals the 0.01 s-sampled one
        ref = po.sampled_maximum_magnitude(s.coeffs, prob["times"], der)
        assert r["max_value"][0] >= ref - 1e-12 and r["max_value"][0] == pytest.approx(ref, abs=0.01)


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_random_batch(po, layout):
    B = 512
    pos, times = random_problems(po, B, 10, 3, seed0=5000)
    coeffs, _ = po.solve_canonical_batch(pos, times, n_threads=8)
    for der in (1, 2):
        r = gpu_extrema(coeffs, times, der, layout=layout)
        check_against_oracle(po, coeffs, times, der, r, range(0, B, 8))
    rh = gpu_extrema(coeffs[:100], times[:100], 2, layout=layout, device=False)
    assert np.array_equal(rh["max_value"], r["max_value"][:100]) and np.array_equal(rh["max_seg"], r["max_seg"][:100])


def test_bounds_the_sampled_sweep(po):
    """Size-independent property at scale: the analytic maximum bounds the 1000-sample sweep of
    mtg_feasibility_batch from above and is reached by it to O(dt^2)."""
    B = 16384
    pos, times = random_problems(po, 64, 10, 3, seed0=800)
    rng = np.random.RandomState(0)
    pos = np.repeat(pos, B // 64, axis=0) + rng.normal(0, 0.5, size=(B, 11, 3))
    times = np.repeat(times, B // 64, axis=0) * rng.uniform(0.8, 1.5, size=(B, 10))
    c = ctx()
    p, t = dev(soa(pos)), dev(soa(times))
    sol = c.solve_batch(p, t)
    tm = c.max_time_batch(t)
    sw = c.feasibility_batch(sol["coeffs"], t, 0.0, tm, tm / 1000, 3.0, 5.0, max_samples=1010, want_flags=False)
    for der, key in ((1, "max_v"), (2, "max_a")):
        ex = c.extrema_batch(sol["coeffs"], t, der)
        an, sa = host(ex["max_value"]), host(sw[key])
        assert np.all(host(ex["status"]) == 0)
        assert np.all(an >= sa * (1 - 1e-12))
        assert np.all(an <= sa * 1.002 + 1e-9)


def test_one_dimension_and_random_polynomials(po):
    """D = 1 takes the roots of p^(d+1) directly (segment.cpp:124-131); random polynomials in the
    spirit of PolynomialTest.FindMinMax (TEST_POLY:81-137): up to 12 coefficients in +-100."""
    rng = np.random.RandomState(1234567)
    B, K = 300, 1
    for n_coef in (10, 12):
        coeffs = rng.uniform(-100, 100, size=(B, K, 1, n_coef))
        coeffs[:, :, :, rng.randint(3, n_coef)] = 0.0          # a zero coefficient somewhere
        coeffs[::7, :, :, -1] = 0.0                            # zero leading coefficient (findLastNonZeroCoeff)
        times = rng.uniform(0.5, 3.0, size=(B, K))
        c = ctx()
        for der in (0, 1):
            r = c.extrema_batch(dev(soa(coeffs)), dev(soa(times)), der)
            mv, mt, Mv, Mt = (host(r[k]) for k in ("min_value", "min_time", "max_value", "max_time"))
            for b in range(B):
                (omt, omv, _), (oMt, oMv, _) = po.traj_min_max_magnitude(coeffs[b], times[b], der)
                assert abs(Mv[b] - oMv) <= 1e-9 * oMv and abs(mv[b] - omv) <= 1e-9 * oMv
                assert abs(Mt[b] - oMt) <= 1e-6 * times[b, 0]


def test_argument_errors():
    import mav_tube_trajectory_generation_b200 as m

    c = ctx()
    coeffs = dev(np.zeros((1, 3, 10, 4)))
    t = dev(np.ones((1, 4)))
    with pytest.raises(m.MtgError):   # LIN_I:400-401 CHECK(N - derivative - 1 > 0)
        c.extrema_batch(coeffs, t, 9)


# ------------------------------------------------------------------ E6 candidate lists, R1 root lists, E5 soft form
def test_candidate_lists_vs_reference_rpoly(po):
    """computeSegmentMaximumMagnitudeCandidates (LIN_I:396-417) = Segment::computeMinMaxMagnitudeCandidateTimes
    (segment.cpp:82-133): [t_start, t_end, real roots in range]. The oracle's roots come from the reference's
    own Jenkins-Traub; the lists are compared as sets on the INTERIOR segments (the rest-to-rest end segments
    carry numerically multiple roots whose list is solver dependent): same count, same
    times to 1e-7 T, same values to 1e-9 of the maximum."""
    B, K = 64, 10
    pos, times = random_problems(po, B, K, 3, seed0=9100)
    coeffs, _ = po.solve_canonical_batch(pos, times, n_threads=8)
    c = ctx()
    for layout in ("soa", "aos"):
        conv_in = soa if layout == "soa" else np.ascontiguousarray
        conv = aos if layout == "soa" else (lambda x: x)
        for der in (1, 2):
            r = c.extrema_candidates_batch(dev(conv_in(coeffs)), dev(conv_in(times)), der, layout=layout)
            ct, cv, nc = conv(host(r["cand_time"])), conv(host(r["cand_value"])), conv(host(r["n_candidates"]))
            assert np.all(host(r["status"]) == 0)
            ex = c.extrema_batch(dev(conv_in(coeffs)), dev(conv_in(times)), der, layout=layout, want_segments=True)
            seg_max = conv(host(ex["seg_max_value"]))
            n_cmp = 0
            for b in range(B):
                for s in range(K):
                    n = nc[b, s]
                    assert n >= 2 and ct[b, s, 0] == 0.0 and ct[b, s, 1] == times[b, s]
                    assert np.all(np.diff(ct[b, s, 2:n]) > 0)                       # roots ascending
                    assert cv[b, s, :n].max() == seg_max[b, s]                      # the list IS what the maximum is taken over
                    vals = [np.sqrt(sum(po.poly_evaluate(coeffs[b, s, d_], t, der) ** 2 for d_ in range(3)))
                            for t in ct[b, s, :n]]
                    assert np.allclose(cv[b, s, :n], vals, rtol=1e-12, atol=1e-12 * seg_max[b].max())
                    if s in (0, K - 1):
                        continue
                    want = po.segment_candidate_times(coeffs[b, s], der, 0.0, times[b, s])
                    if len(want) != n:
                        continue       # an even-multiplicity / near-real complex pair on one side only: values are checked above
                    n_cmp += 1
                    assert np.abs(np.sort(want[2:]) - ct[b, s, 2:n]).max(initial=0.0) <= 1e-7 * times[b, s]
            assert n_cmp > 0.9 * B * (K - 2)


def test_candidate_interval_and_dimension_subset(po):
    """t_start / t_end and `dimensions` of Segment::computeMinMaxMagnitudeCandidateTimes (segment.cpp:82-133):
    a sub-interval of every segment, all dimensions and a single one (roots of p^(d+1), :124-131)."""
    B, K = 32, 4
    pos, times = random_problems(po, B, K, 3, seed0=9300)
    coeffs, _ = po.solve_canonical_batch(pos, times, n_threads=8)
    lo, hi = 0.25 * times, 0.8 * times
    c = ctx()
    for mask, dims in ((0, None), (0b010, [1]), (0b101, [0, 2])):
        r = c.extrema_candidates_batch(dev(soa(coeffs)), dev(soa(times)), 1, t_start=dev(soa(lo)), t_end=dev(soa(hi)),
                                       dim_mask=mask)
        ct, cv, nc = aos(host(r["cand_time"])), aos(host(r["cand_value"])), aos(host(r["n_candidates"]))
        n_cmp = 0
        for b in range(B):
            for s in range(K):
                n = nc[b, s]
                assert ct[b, s, 0] == lo[b, s] and ct[b, s, 1] == hi[b, s]
                assert np.all((ct[b, s, :n] >= lo[b, s]) & (ct[b, s, :n] <= hi[b, s]))
                want = po.segment_candidate_times(coeffs[b, s], 1, lo[b, s], hi[b, s], dims=dims)
                dd = dims if dims is not None else [0, 1, 2]
                vals = [np.sqrt(sum(po.poly_evaluate(coeffs[b, s, d_], t, 1) ** 2 for d_ in dd)) for t in ct[b, s, :n]]
                assert np.allclose(cv[b, s, :n], vals, rtol=1e-12, atol=1e-13)
                if len(want) == n:
                    n_cmp += 1
                    assert np.abs(np.sort(want[2:]) - ct[b, s, 2:n]).max(initial=0.0) <= 1e-7 * times[b, s]
        assert n_cmp > 0.9 * B * K


@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_real_roots_vs_jenkins_traub(po, layout):
    """findRootsJenkinsTraub (rpoly_ak1.cpp:70-117) + the real / in-range selection of polynomial.cpp:46-60 on
    random polynomials (simple roots): same count, same roots to 1e-9, for every degree up to 21."""
    rng = np.random.RandomState(99)
    c = ctx()
    for n in (2, 3, 6, 11, 16, 22):
        B = 200
        coeffs = rng.uniform(-1, 1, size=(B, n))
        coeffs[::9, -1] = 0.0                                     # zero leading coefficient
        lo, hi = rng.uniform(-3, -0.5, size=B), rng.uniform(0.5, 3, size=B)
        cc = coeffs if layout == "aos" else soa(coeffs)
        r = c.poly_real_roots_batch(dev(cc), dev(lo), dev(hi), layout=layout)
        roots = host(r["roots"]) if layout == "aos" else aos(host(r["roots"]))
        nr = host(r["n_roots"])
        assert np.all(host(r["status"]) == 0)
        n_ok = 0
        for b in range(B):
            ok, z = po.find_roots_jenkins_traub(coeffs[b])
            want = np.sort(np.array([x.real for x in z if abs(x.imag) <= np.finfo(float).eps
                                     and lo[b] <= x.real <= hi[b]]))
            got = roots[b, :nr[b]]
            if len(want) != nr[b]:
                continue                                           # a (near-)double root: counted below
            n_ok += 1
            assert np.abs(got - want).max(initial=0.0) <= 1e-9 * max(1.0, np.abs(want).max(initial=0.0))
        assert n_ok >= 0.97 * B
    # host-memory mode and an exact root on the boundary: t^2 - 1 on [-1, 1], t (t - 0.5) on [0, 1]
    cc = np.array([[-1.0, 0.0, 1.0], [0.0, -0.5, 1.0]])
    r = c.poly_real_roots_batch(cc, np.array([-1.0, 0.0]), np.array([1.0, 1.0]), layout="aos")
    assert list(r["n_roots"]) == [2, 2]
    assert np.allclose(r["roots"][0, :2], [-1.0, 1.0], atol=1e-15) and np.allclose(r["roots"][1, :2], [0.0, 0.5], atol=1e-15)


def test_soft_constraint(po):
    """evaluateMaximumMagnitudeAsSoftConstraint (NL_I:2735-2766): sum over (derivative, limit) of
    min(max_cost, exp((max - limit) / limit * weight)), max from computeMaximumOfMagnitude (LIN_I:455-487)."""
    B = 96
    pos, times = random_problems(po, B, 10, 3, seed0=6200)
    coeffs, _ = po.solve_canonical_batch(pos, times, n_threads=8)
    c = ctx()
    ders, lims, w, cap = [1, 2], [3.0, 5.0], 100.0, 1e12
    r = c.soft_constraint_batch(dev(soa(coeffs)), dev(soa(times)), ders, lims, w, cap)
    cost, viol = host(r["cost"]), host(r["violations"])
    assert np.all(host(r["status"]) == 0)
    for b in range(B):
        want = 0.0
        for q, (d_, lim) in enumerate(zip(ders, lims)):
            _, v, _ = po.opt_max_magnitude(coeffs[b], times[b], d_)
            assert abs(viol[q, b] - (v - lim)) <= 1e-9 * v
            want += min(cap, np.exp((v - lim) / lim * w))
        # the exponential amplifies the 1e-9 value tolerance by weight / limit
        assert abs(cost[b] - want) <= 1e-6 * want


@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_two_isolators_agree_at_scale(po, layout):
    """Size-independent property: the kernel has two independent isolators — the compile-time path of the default
    problem (lane = interval, full degree) when no interval is given, and the generic path (16 lanes per
    interval, per-problem degree, two-sided) when t_start / t_end are explicit. On 8,192 trajectories x 10
    segments x {velocity, acceleration} both must produce the same candidate lists: same count on all but a
    few near-degenerate segments, times to 1e-9 T, maxima to 1e-12."""
    import torch
    B, K = 8192, 10
    c = ctx()
    pos, t = c.generate_candidates_batch(B, K, 3, seed=77, first_index=0, pos_min=[-10.0] * 3, pos_max=[10.0] * 3,
                                         v_max=3.0, a_max=5.0, layout=layout)
    sol = c.solve_batch(pos, t, layout=layout)
    conv = aos if layout == "soa" else (lambda x: x)
    for der in (1, 2):
        fast = c.extrema_candidates_batch(sol["coeffs"], t, der, layout=layout)
        gen = c.extrema_candidates_batch(sol["coeffs"], t, der, t_start=torch.zeros_like(t), t_end=t, layout=layout)
        assert int((fast["status"] != 0).sum()) == 0 and int((gen["status"] != 0).sum()) == 0
        nf, ng = conv(host(fast["n_candidates"])), conv(host(gen["n_candidates"]))
        tf, tg = conv(host(fast["cand_time"])), conv(host(gen["cand_time"]))
        vf, vg = conv(host(fast["cand_value"])), conv(host(gen["cand_value"]))
        T = conv(host(t))
        same = nf == ng
        assert same.mean() > 0.999
        col = np.arange(tf.shape[2])[None, None, :]
        live = same[:, :, None] & (col < nf[:, :, None])
        assert np.abs(np.where(live, tf - tg, 0.0)).max() <= 1e-9 * T.max()
        # the maximum over the list is what computeMaximumOfMagnitude consumes: equal on EVERY segment
        mf = np.where(col < nf[:, :, None], vf, -np.inf).max(axis=2)
        mg = np.where(col < ng[:, :, None], vg, -np.inf).max(axis=2)
        assert np.abs(mf - mg).max() <= 1e-12 * mf.max()

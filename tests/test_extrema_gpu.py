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
            assert r["min_seg"][b] == ms and abs(r["min_time"][b] - mt) <= 1e-6 * times[b][ms]
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

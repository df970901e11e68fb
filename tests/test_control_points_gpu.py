"""-m gpu parity tests of mtg_control_points_batch (N2: Bezier control points + the reference's tube / end-cap /
sphere constraints evaluated on them, QC_I:267-474) against the oracle, and the property that re-pins the
SAMPLED tube predicate of mtg_feasibility_batch (T1) to something derived from the reference other than
itself: a polynomial lies in the convex hull of its control points and the tube-and-caps region is convex, so
all control points feasible  ==>  every sample of the sweep carries the in-tube bit."""
import math

import numpy as np
import pytest

from gpu_util import aos, ctx, dev, host, random_problems, soa

pytestmark = pytest.mark.gpu
N, H = 10, 5


def endpoint_derivatives(po, coeffs, times):
    """[B,K,D,N], [B,K] -> C [d_f; d_p] as [B,K+1,h,D] (start of every segment + end of the last)."""
    B, K, D, _ = coeffs.shape
    der = np.zeros((B, K + 1, H, D))
    for b in range(B):
        for v in range(K + 1):
            seg, t = (v, 0.0) if v < K else (K - 1, times[b, K - 1])
            for k in range(H):
                for dim in range(D):
                    der[b, v, k, dim] = po.poly_evaluate(coeffs[b, seg, dim], t, k)
    return der


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_control_points_and_constraints_vs_oracle(po, layout):
    B, K = 48, 10
    pos, times = random_problems(po, B, K, 3, seed0=4100)
    coeffs, _ = po.solve_canonical_batch(pos, times, n_threads=8)
    der = endpoint_derivatives(po, coeffs, times)
    radii = np.random.RandomState(2).uniform(0.2, 3.0, size=(B, K, 2))
    conv_in = soa if layout == "soa" else np.ascontiguousarray
    conv = aos if layout == "soa" else (lambda x: x)
    c = ctx()
    for use_der in (True, False):
        r = c.control_points_batch(dev(conv_in(times)), coeffs=None if use_der else dev(conv_in(coeffs)),
                                   derivatives=dev(conv_in(der)) if use_der else None, positions=dev(conv_in(pos)),
                                   radii=dev(conv_in(radii)), layout=layout)
        cps, tube, cs, ce, sph = (conv(host(r[k])) for k in ("control_points", "tube", "cap_start", "cap_end", "sphere"))
        mx, fe = host(r["max_value"]), host(r["feasible"])
        assert np.all(host(r["status"]) == 0)
        # from the coefficients the endpoint derivatives are re-evaluated: same values to rounding
        tol = 1e-12 if use_der else 1e-9
        for b in range(B):
            o = po.control_point_constraints(der[b], times[b], pos[b], radii[b])
            scale = np.abs(o["control_points"]).max()
            assert np.abs(cps[b] - o["control_points"]).max() <= tol * scale
            assert np.abs(tube[b] - o["tube"]).max() <= 10 * tol * scale ** 2
            assert np.abs(cs[b] - o["cap_start"]).max() <= 10 * tol * scale
            assert np.abs(ce[b] - o["cap_end"]).max() <= 10 * tol * scale
            assert np.abs(sph[b, :-1] - o["sphere"][:-1]).max() <= 10 * tol * scale ** 2 and sph[b, -1] == -np.inf
            want_max = max(o["tube"].max(), o["cap_start"].max(), o["cap_end"].max(), o["sphere"].max())
            assert abs(mx[b] - want_max) <= 10 * tol * scale ** 2
            if abs(want_max) > 1e-6:
                assert bool(fe[b]) == (want_max <= 0.0)
    # host-memory mode == device mode, bit for bit
    rh = c.control_points_batch(conv_in(times), coeffs=conv_in(coeffs), positions=conv_in(pos), radii=conv_in(radii),
                                layout=layout)
    assert np.array_equal(conv(rh["control_points"]), cps) and np.array_equal(conv(rh["tube"]), tube)
    assert np.array_equal(rh["feasible"], fe)


def test_control_points_reproduce_the_polynomial(po):
    """Bernstein form: sum_j cp_j binom(n, j) u^j (1-u)^(n-j) = p(u T) — the identity that pins the mapping of
    QC_I:267-319 (any N, D; here N = 10, 8 and D = 3, 1)."""
    c = ctx()
    for n_coef, D, K in ((10, 3, 6), (8, 1, 3)):
        rng = np.random.RandomState(n_coef)
        B = 32
        coeffs = rng.uniform(-1, 1, size=(B, K, D, n_coef))
        times = rng.uniform(1.0, 4.0, size=(B, K))
        r = c.control_points_batch(dev(soa(times)), coeffs=dev(soa(coeffs)), N=n_coef)
        cps = aos(host(r["control_points"]))                       # [B,K,N,D]
        n = n_coef - 1
        for u in (0.0, 0.3, 0.77, 1.0):
            bern = np.array([math.comb(n, j) * u ** j * (1 - u) ** (n - j) for j in range(n_coef)])
            x = np.einsum("j,bkjd->bkd", bern, cps)
            want = np.einsum("bkdj,bkj->bkd", coeffs, (u * times)[..., None] ** np.arange(n_coef))
            assert np.abs(x - want).max() <= 1e-9 * np.abs(want).max()


def test_control_points_inside_imply_every_sample_inside(po):
    """T1 re-pin. Trajectories whose control points all satisfy the reference's tube + end-cap constraints
    (QC_I:369-474) must have the in-tube bit on EVERY sample of mtg_feasibility_batch (and of the oracle's
    sampled predicate): convex hull + convex region. Radii are chosen so that a good share of the batch is
    control-point feasible; the others show that the sampled predicate does reject."""
    B, K = 2048, 10
    pos, times = random_problems(po, 64, K, 3, seed0=8800)
    rng = np.random.RandomState(3)
    pos = np.repeat(pos, B // 64, axis=0) + rng.normal(0, 0.2, size=(B, K + 1, 3))
    times = np.repeat(times, B // 64, axis=0) * rng.uniform(0.9, 1.3, size=(B, K))
    radii = np.empty((B, K, 2))
    radii[..., 0] = rng.uniform(0.5, 25.0, size=(B, 1))       # tube radius: from tight to generous
    radii[..., 1] = rng.uniform(0.5, 25.0, size=(B, 1))       # cap / sphere radius
    c = ctx()
    p, t, rd = dev(soa(pos)), dev(soa(times)), dev(soa(radii))
    sol = c.solve_batch(p, t)
    cp = c.control_points_batch(t, coeffs=sol["coeffs"], positions=p, radii=rd)
    tube_ok = ((host(cp["tube"]) <= 0) & (host(cp["cap_start"]) <= 0) & (host(cp["cap_end"]) <= 0)).all(axis=(0, 1))
    tm = c.max_time_batch(t)
    sw = c.feasibility_batch(sol["coeffs"], t, 0.0, tm, tm / 1000, 3.0, 5.0, positions=p, radii=rd, max_samples=1010)
    flags, n = aos(host(sw["flags"])), host(sw["n_samples"])
    all_in = np.array([np.all(flags[b, :n[b]] & 4) for b in range(B)])
    assert tube_ok.sum() > B // 20 and (~tube_ok).sum() > B // 20
    assert np.all(all_in[tube_ok]), np.flatnonzero(tube_ok & ~all_in)[:5]
    assert (~all_in).sum() > 0          # the sampled predicate does reject somewhere
    # the oracle's sampled predicate agrees on a subset (it is the checker of the flags elsewhere)
    coeffs = aos(host(sol["coeffs"]))
    for b in np.flatnonzero(tube_ok)[:16]:
        ref = po.feasibility_sweep(coeffs[b], times[b], pos[b], radii[b], 3.0, 5.0, 0.0, host(tm)[b], host(tm)[b] / 1000)
        assert np.all(ref[1] & 4)

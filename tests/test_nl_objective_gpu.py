"""-m gpu parity tests of the batched non-linear objective (N1) against the oracle's restatement of
getCostAndGradientDerivative (NL_I:1537-1606), evaluateMaximumMagnitudeAsSoftConstraint (NL_I:2735-2766) and its
finite-difference gradients (NL_I:2365-2490), plus the open-loop check of the projected-gradient driver: at every
iterate the recorded J_d / J_sc equal the oracle's values for that iterate's free derivatives.

Tolerances: J_d 1e-9 against the 60-digit-pinned cost (the oracle's dense d^T R d on its rounding-noisy R is itself
only ~5e-8 accurate, see test_cost_fd_gpu.py): 1e-7 against the oracle. The analytic gradient 2 R_pf d_f + 2 R_pp d_p
cancels to ~0 at the optimum, so it is compared relative to the size of its two terms. FD gradients of the soft
cost: 1e-7 of the cost scale (the north-star's bar for FD gradients)."""
import numpy as np
import pytest

from gpu_util import aos, ctx, dev, host, random_problems, soa

pytestmark = pytest.mark.gpu
N, NF = 10, 4


def solved(po, B, K, seed0):
    pos, times = random_problems(po, B, K, 3, seed0=seed0)
    c = ctx()
    r = c.solve_batch(dev(np.ascontiguousarray(pos)), dev(np.ascontiguousarray(times)), want_free=True, layout="aos")
    return pos, times, host(r["free"]), host(r["coeffs"])


@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_cost_and_gradient_derivative(po, layout):
    B, K = 64, 10
    pos, times, free, _ = solved(po, B, K, 3100)
    rng = np.random.RandomState(0)
    free = free + rng.normal(0, 0.3, size=free.shape)           # away from the optimum: a gradient that is not ~0
    conv_in = soa if layout == "soa" else np.ascontiguousarray
    conv = aos if layout == "soa" else (lambda x: x)
    c = ctx()
    r = c.cost_derivative_batch(dev(conv_in(pos)), dev(conv_in(times)), dev(conv_in(free)), layout=layout, want_diag=True)
    J, g, dg = host(r["J_d"]), conv(host(r["grad"])), conv(host(r["diag"]))
    for b in range(B):
        mask, values = po.canonical_mask_values(pos[b])
        s = po.solve(N, 4, times[b], mask, values)
        Jo, go = po.cost_gradient_derivative(N, 4, times[b], mask, values, free[b].reshape(3, -1))
        assert abs(J[b] - Jo) <= 1e-7 * Jo
        Rpp = s.R[s.n_fixed:, s.n_fixed:]
        scale = 2 * np.abs(Rpp) @ np.abs(free[b].reshape(3, -1)).T        # size of the terms that cancel
        assert np.all(np.abs(g[b].reshape(3, -1) - go) <= 1e-7 * scale.T + 1e-12)
        assert np.allclose(dg[b].reshape(-1), 2 * np.diag(Rpp), rtol=1e-7)
    # consistency with the solve: at the optimum the gradient vanishes relative to its terms, and J_d = 2 * cost
    pos, times, free, _ = solved(po, B, K, 3100)
    sol = c.solve_batch(dev(conv_in(pos)), dev(conv_in(times)), layout=layout)
    r = c.cost_derivative_batch(dev(conv_in(pos)), dev(conv_in(times)), dev(conv_in(free)), layout=layout)
    assert np.allclose(host(r["J_d"]), 2 * host(sol["cost"]), rtol=1e-10)
    # translation of d_p along the gradient lowers nothing: first-order optimality, |g| tiny against J_d / |d_p|
    g = conv(host(r["grad"]))
    assert np.abs(g).max() <= 1e-6 * host(r["J_d"]).max()


def test_soft_constraint_gradient(po):
    B, K = 24, 6
    pos, times, free, coeffs = solved(po, B, K, 3300)
    ders, lims, w, cap, inc = [1, 2], [2.0, 2.0], 10.0, 1e12, 0.05
    c = ctx()
    for central in (True, False):
        r = c.soft_constraint_gradient_batch(dev(coeffs), dev(np.ascontiguousarray(times)), ders, lims, w, cap, inc,
                                             central=central)
        J, g = host(r["J_sc"]), host(r["grad"])
        assert np.all(host(r["status"]) == 0)
        for b in range(0, B, 3):
            mask, values = po.canonical_mask_values(pos[b])
            Jo, go = po.soft_constraint_gradient(N, 4, times[b], mask, values, free[b].reshape(3, -1), ders, lims, w, cap,
                                                 inc, central=central)
            assert abs(J[b] - Jo) <= 1e-7 * Jo
            # both sides difference two costs that agree to ~1e-9 J (extremum values 1e-9, amplified by weight / limit)
            assert np.abs(g[b].reshape(3, -1) - go).max() <= 1e-7 * Jo / inc + 1e-7 * np.abs(go).max()


def test_descent_driver_open_loop(po):
    """mtg_nl_descent_batch. Open-loop parity: the J_d / J_sc recorded for the RETURNED point are the oracle's values
    for the returned d_p, whose coefficients are the oracle's setFreeConstraints coefficients. Driver properties:
    the bounds of NL_I:2858-2905 hold, the returned point never has a larger objective than the start, the
    accepted trial points are non-increasing, and splitting a run into single steps reproduces it bit for bit."""
    import torch

    B, K, iters = 32, 6, 8
    pos, times, _, _ = solved(po, B, K, 3500)
    times = times * 0.6                                   # too fast: v / a limits are violated, the soft term pushes
    c = ctx()
    p, t = dev(np.ascontiguousarray(pos)), dev(np.ascontiguousarray(times))
    sol = c.solve_batch(p, t, want_free=True, layout="aos")
    ders, lims = [1, 2], [3.0, 5.0]
    kw = dict(w_d=1.0, w_sc=1.0, soft_weight=5.0, increment=0.05, step=0.5)
    x = sol["free"].clone()
    r = c.nl_descent_batch(p, t, x, ders, lims, iterations=iters, **kw)
    h = host(r["history"])                                 # [iters + 2, 3, B]
    assert np.all(host(r["status"]) == 0)
    f = h[:, 0, :] + h[:, 1, :]
    acc = h[:, 2, :] > 0.5
    assert np.all(acc[0]) and np.all(acc[-1])
    assert np.all(f[-1] <= f[0])                           # never worse than the start
    assert np.any(f[-1] < 0.999 * f[0])                    # and it does make progress somewhere
    for b in range(B):                                     # accepted trial points are non-increasing
        fa = f[:-1, b][acc[:-1, b]]
        assert np.all(np.diff(fa) <= 0.0)
        assert f[-1, b] == fa[-1]                          # the returned point is the last accepted one
    xf = host(x)                                           # [B, D, K-1, NF]
    assert np.all(np.abs(xf[..., 0]) <= 3.0 + 1e-12) and np.all(np.abs(xf[..., 1]) <= 5.0 + 1e-12)   # NL_I:2858-2905
    # zero iterations = evaluate only: the start is returned unchanged
    x0 = sol["free"].clone()
    r0 = c.nl_descent_batch(p, t, x0, ders, lims, iterations=0, **kw)
    start = sol["free"].clone()                           # the start point is projected onto the bounds first
    start[..., 0].clamp_(-3.0, 3.0)
    start[..., 1].clamp_(-5.0, 5.0)
    assert torch.equal(x0, start) and np.array_equal(host(r0["history"])[0, :2], h[0, :2])
    # open loop: the returned point against the oracle
    cf = host(r["coeffs"])
    for b in range(0, B, 4):
        mask, values = po.canonical_mask_values(pos[b])
        Jo, _ = po.cost_gradient_derivative(N, 4, times[b], mask, values, xf[b].reshape(3, -1))
        Jso, _ = po.soft_constraint_gradient(N, 4, times[b], mask, values, xf[b].reshape(3, -1), ders, lims, 5.0, 1e12,
                                             0.05, want_grad=False)
        assert abs(h[-1, 0, b] - Jo) <= 1e-7 * Jo and abs(h[-1, 1, b] - Jso) <= 1e-7 * Jso
        want_c = po.coeffs_from_free_constraints(N, times[b], mask, values, xf[b].reshape(3, -1))
        den = np.abs(want_c).max(axis=-1)
        assert (np.abs(cf[b] - want_c).max(axis=-1) / den).max() <= 1e-9
    # the whole descent is capturable in a CUDA graph (the library only enqueues work on the caller's stream)
    xg = sol["free"].clone()
    g = torch.cuda.CUDAGraph()
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        c.nl_descent_batch(p, t, xg.clone(), ders, lims, iterations=1, **kw)     # work spaces allocated outside the capture
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=stream):
            rg = c.nl_descent_batch(p, t, xg, ders, lims, iterations=iters, **kw)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(xg, x) and np.array_equal(host(rg["history"]), h)

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
    """Every iterate of mtg_nl_descent_batch: the recorded J_d / J_sc are the oracle's for that iterate's d_p
    (open-loop parity), the iterates respect the bounds, and the weighted objective does not increase."""
    import torch

    B, K, iters = 32, 6, 6
    pos, times, free0, _ = solved(po, B, K, 3500)
    times = times * 0.6                                   # too fast: v / a limits are violated, the soft term pushes
    c = ctx()
    sol = c.solve_batch(dev(np.ascontiguousarray(pos)), dev(np.ascontiguousarray(times)), want_free=True, layout="aos")
    ders, lims = [1, 2], [3.0, 5.0]
    p, t = dev(np.ascontiguousarray(pos)), dev(np.ascontiguousarray(times))
    frees = []
    x = sol["free"].clone()
    # iterate one step at a time to capture every d_p (the driver itself records only the costs)
    hist = []
    for it in range(iters):
        r = c.nl_descent_batch(p, t, x, ders, lims, w_d=1.0, w_sc=1.0, soft_weight=5.0, increment=0.05, step=0.3,
                               iterations=1)
        hist.append(host(r["history"])[0])
        frees.append(host(x).copy() if it == iters - 1 else None)
    # one multi-iteration call from the same start gives the same history (and is what a user runs)
    x2 = sol["free"].clone()
    r2 = c.nl_descent_batch(p, t, x2, ders, lims, w_d=1.0, w_sc=1.0, soft_weight=5.0, increment=0.05, step=0.3,
                            iterations=iters)
    h2 = host(r2["history"])
    assert np.allclose(h2[:iters], np.stack(hist), rtol=1e-12, atol=0)
    assert torch.equal(x, x2)
    xf = host(x2)                                          # [B, D, K-1, NF]
    assert np.all(np.abs(xf[..., 0]) <= 3.0 + 1e-12) and np.all(np.abs(xf[..., 1]) <= 5.0 + 1e-12)   # NL_I:2858-2905
    total = h2[:, 0, :] + h2[:, 1, :]
    assert np.all(total[-1] <= total[0] * (1 + 1e-9))
    # open loop: the final point against the oracle
    cf = host(r2["coeffs"])
    for b in range(0, B, 4):
        mask, values = po.canonical_mask_values(pos[b])
        Jo, _ = po.cost_gradient_derivative(N, 4, times[b], mask, values, xf[b].reshape(3, -1))
        Jso, _ = po.soft_constraint_gradient(N, 4, times[b], mask, values, xf[b].reshape(3, -1), ders, lims, 5.0, 1e12,
                                             0.05, want_grad=False)
        assert abs(h2[-1, 0, b] - Jo) <= 1e-7 * Jo and abs(h2[-1, 1, b] - Jso) <= 1e-7 * Jso
        want_c = po.coeffs_from_free_constraints(N, times[b], mask, values, xf[b].reshape(3, -1))
        den = np.abs(want_c).max(axis=-1)
        assert (np.abs(cf[b] - want_c).max(axis=-1) / den).max() <= 1e-9

"""-m gpu parity tests of mtg_cost_time_fd_batch (P9: the finite-difference segment-time
perturbation batch of the reference's non-linear layer, NL_I:2495-2657 + 1537-1606).

Bars: the nominal and perturbed costs J_d are compared (a) with the oracle, which restates the
reference's evaluation order (dense d^T R d on a rounding-noisy R: itself only ~5e-8 from the
exact value, tests/test_oracle.py::test_cost_time_fd) at 1e-7 relative, and (b) with a 60-digit
evaluation of the same formulas at 1e-12 relative (the 1e-9 bar of BASELINE.json with margin).
Gradients are differences of nearly equal numbers: their error bound eps*|q|/delta is written
in the test.
"""
import numpy as np
import pytest

from conftest import REFERENCE_PARAMS, make_reference_problem
from gpu_util import aos, ctx, dev, host, random_problems, soa

pytestmark = pytest.mark.gpu
N = 10


def gpu_fd(pos, times, free, inc, central, derivative=4, layout="soa", device=True):
    """pos [B,K+1,D], times [B,K], free [B,D,K-1,4] -> J [B], Jp/Jm/grad [B,K]."""
    c = ctx()
    conv_in = soa if layout == "soa" else np.ascontiguousarray
    p, t, f = conv_in(pos), conv_in(times), conv_in(free)
    if device:
        p, t, f = dev(p), dev(t), dev(f)
    r = c.cost_time_fd_batch(p, t, f, inc, central=central, N=N, derivative=derivative, layout=layout)
    if device:
        import torch

        torch.cuda.synchronize()
    conv = aos if layout == "soa" else (lambda x: x)
    return (host(r["J"]), conv(host(r["J_plus"])), conv(host(r["J_minus"])) if central else None,
            conv(host(r["grad"])), host(r["status"]))


def solve_free(pos, times, derivative=4):
    r = ctx().solve_batch(dev(soa(pos)), dev(soa(times)), N=N, derivative=derivative, want_free=True)
    return aos(host(r["free"])), host(r["cost"])


@pytest.mark.parametrize("name", ["segment_10_dim_3", "segment_10_dim_1", "segment_50_dim_3", "accel_5_dim_3",
                                  "jerk_5_dim_3", "segment_1_dim_3"])
@pytest.mark.parametrize("inc", [1e-3, 1e-6])
def test_reference_parameter_sets(po, name, inc):
    from exact_solver import exact_cost_derivative

    prob = make_reference_problem(name)
    der, K, D = prob["derivative"], prob["K"], prob["D"]
    pos = prob["values"][None, :, 0, :]
    times = prob["times"][None]
    s = po.solve(N, der, prob["times"], prob["mask"], prob["values"])
    d_p = s.d_p.reshape(D, max(K - 1, 0), 4)
    J, Jp, Jm, g, st = gpu_fd(pos, times, d_p[None], inc, True, der)
    assert st[0] == 0
    J0, oJp, oJm, og = po.cost_time_fd(N, der, prob["times"], prob["mask"], prob["values"], s.d_p, inc, True)
    # (a) oracle = reference evaluation order
    assert abs(J[0] - J0) <= 1e-7 * J0
    assert np.all(np.abs(Jp[0] - oJp) <= 1e-7 * J0) and np.all(np.abs(Jm[0] - oJm) <= 1e-7 * J0)
    # (b) 60-digit evaluation of the same formulas
    Je = exact_cost_derivative(N, der, prob["times"], prob["mask"], prob["values"], s.d_p.reshape(D, -1))
    assert abs(J[0] - float(Je)) <= 1e-12 * float(Je)
    for n in (0, K // 2, K - 1):
        tp, tm = prob["times"].copy(), prob["times"].copy()
        tp[n] = 0.1 if tp[n] <= 0.1 else tp[n] + inc
        tm[n] = 0.1 if tm[n] <= 0.1 else tm[n] - inc
        Jpe = exact_cost_derivative(N, der, tp, prob["mask"], prob["values"], s.d_p.reshape(D, -1))
        Jme = exact_cost_derivative(N, der, tm, prob["mask"], prob["values"], s.d_p.reshape(D, -1))
        assert abs(Jp[0, n] - float(Jpe)) <= 1e-12 * float(Je)
        assert abs(Jm[0, n] - float(Jme)) <= 1e-12 * float(Je)
        ge = float((Jpe - Jme) / (2 * inc))
        # rounding of the two segment terms (<= J each) divided by 2*inc
        assert abs(g[0, n] - ge) <= 64 * 2.2e-16 * float(Je) / inc + 1e-12 * abs(ge)
    # forward differences (getCostAndGradientTimeSimple)
    Jf, Jpf, _, gf, _ = gpu_fd(pos, times, d_p[None], inc, False, der)
    assert Jf[0] == J[0] and np.array_equal(Jpf, Jp)
    assert np.all(np.abs(gf[0] - (Jp[0] - J[0]) / inc) <= 64 * 2.2e-16 * J[0] / inc)


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_batch_matches_solve_cost_and_oracle(po, layout):
    """4,096 trajectories (BASELINE config 3): J_d(T) = 2 computeCost at the solved d_p (LIN_I:129 vs
    NL_I:1585-1588); a sample of the batch against the oracle; host-memory mode == device mode."""
    B = 4096
    pos, times = random_problems(po, B, 10, 3, seed0=9000)
    free, cost = solve_free(pos, times)
    J, Jp, Jm, g, st = gpu_fd(pos, times, free, 0.1, True, layout=layout)
    assert np.all(st == 0)
    # two evaluation orders of the same sum of squares: |W dhat|^2 here, |Lt chat|^2 in the solve's epilogue
    assert np.all(np.abs(J - 2 * cost) <= 1e-10 * J)
    for b in range(0, B, 512):
        mask, values = po.canonical_mask_values(pos[b])
        J0, oJp, oJm, og = po.cost_time_fd(N, 4, times[b], mask, values, free[b].reshape(-1), 0.1, True)
        assert abs(J[b] - J0) <= 1e-7 * J0
        assert np.all(np.abs(Jp[b] - oJp) <= 1e-7 * J0) and np.all(np.abs(Jm[b] - oJm) <= 1e-7 * J0)
        assert np.all(np.abs(g[b] - og) <= 1e-6 * np.abs(og).max())
    Jh, Jph, Jmh, gh, sth = gpu_fd(pos[:300], times[:300], free[:300], 0.1, True, layout=layout, device=False)
    assert np.array_equal(Jh, J[:300]) and np.array_equal(Jph, Jp[:300]) and np.array_equal(gh, g[:300])


def test_floor_and_errors(po):
    pos, times = random_problems(po, 8, 10, 3, seed0=31)
    free, _ = solve_free(pos, times)
    t2 = times.copy()
    t2[:, 3] = 0.05    # <= 0.1 s: both perturbed times become exactly 0.1 (NL_I:2529-2530)
    t2[:, 5] = 0.1
    J, Jp, Jm, g, st = gpu_fd(pos, t2, free, 1e-3, True)
    assert np.array_equal(Jp[:, 3], Jm[:, 3]) and np.all(g[:, 3] == 0.0)
    assert np.array_equal(Jp[:, 5], Jm[:, 5]) and np.all(g[:, 5] == 0.0)
    mask, values = po.canonical_mask_values(pos[0])
    J0, oJp, oJm, og = po.cost_time_fd(N, 4, t2[0], mask, values, free[0].reshape(-1), 1e-3, True)
    assert np.all(np.abs(Jp[0] - oJp) <= 1e-6 * J0)     # tiny segments: the reference order loses digits
    # non-positive time -> status bit, no abort (reference: CHECK_GT abort, LIN_I:296)
    t3 = times.copy()
    t3[2, 4] = 0.0
    _, _, _, _, st = gpu_fd(pos, t3, free, 1e-3, True)
    assert st[2] == 1 and st[0] == 0
    import mav_tube_trajectory_generation_b200 as m

    with pytest.raises(m.MtgError):
        gpu_fd(pos, times, free, 0.0, True)


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_set_free_constraints_round_trip(po, layout):
    """setFreeConstraints (LIN_I:489-498): feeding the solved d_p back reproduces the solve's coefficients
    bit for bit (same code path) and its cost to rounding, a perturbed d_p matches the oracle's
    coefficients-from-constraints (P8c) and raises the cost (d_p is the minimiser)."""
    c = ctx()
    pos, times = random_problems(po, 256, 10, 3, seed0=600)
    conv_in = soa if layout == "soa" else np.ascontiguousarray
    p, t = dev(conv_in(pos)), dev(conv_in(times))
    sol = c.solve_batch(p, t, want_free=True, layout=layout)
    r = c.set_free_constraints_batch(p, t, sol["free"], layout=layout)
    conv = aos if layout == "soa" else (lambda x: x)
    assert np.array_equal(host(r["coeffs"]), host(sol["coeffs"]))
    assert np.allclose(host(r["cost"]), host(sol["cost"]), rtol=1e-14, atol=0)   # segment sums in another order
    free = conv(host(sol["free"])).copy()
    free2 = free * (1.0 + 0.05 * np.random.RandomState(2).normal(size=free.shape))
    r2 = c.set_free_constraints_batch(p, t, dev(conv_in(free2)), layout=layout)
    assert np.all(host(r2["cost"]) > host(sol["cost"]))
    c2 = conv(host(r2["coeffs"]))
    for b in range(0, 256, 32):
        mask, values = po.canonical_mask_values(pos[b])
        ref = po.coeffs_from_free_constraints(N, times[b], mask, values, free2[b].reshape(-1))
        den = np.abs(ref).max(axis=-1)
        assert (np.abs(c2[b] - ref).max(axis=-1) / den).max() < 1e-9

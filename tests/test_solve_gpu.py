"""-m gpu parity tests of mtg_solve_batch (P1..P8) against the CPU oracle, the
committed golden fixtures and the 60-digit solutions stored with them.

Tolerances (BASELINE.json north_star + SURVEY.md §0.2):
  coefficients : ||c_gpu - c_ref||_inf <= 1e-9 * ||c_ref||_inf per polynomial (segment x dimension)
  cost         : 1e-9 relative
The oracle restates the reference's fp64 evaluation order, which is itself only
1e-11..1.5e-9 from the exact solution (tests/test_oracle.py::test_oracle_vs_mpmath);
where the committed exact solution exists the CUDA path is held to 1e-10 against it.
"""
import os

import numpy as np
import pytest

from conftest import REFERENCE_PARAMS, make_reference_problem
from gpu_util import aos, ctx, dev, host, normwise, random_problems, soa

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N = 10
TOL = 1e-9


# The oracle restates the REFERENCE's fp64 evaluation order (invert A(T), form
# A^-T Q A^-1, dense QR). That order is itself 1e-11 .. 3e-7 away from the exact
# solution depending on the cost derivative and the segment times (measured against
# 60-digit mpmath, see DESIGN.md "parity definition"), so outside config C the 1e-9
# budget cannot be spent against the oracle alone: the worst trajectories of each
# batch are arbitrated by the exact solver, where the CUDA path is held to 1e-10.
ORACLE_NOISE = {4: 1e-8, 3: 1e-8, 2: 5e-8, 1: 5e-6, 0: 1e-4}


def arbitrate(po, pos, times, der, coeffs, cost, ref_c, ref_cost, n=1):
    """EVERY item of the batch against the exact solution of the reference's normal equations — the
    binary128 arbiter oracle/exact128.cpp (pinned to the 60-digit mpmath solve in tests/test_oracle.py) —
    at 1e-10 (coefficients, norm-wise per polynomial) / 1e-11 (cost) for cost derivatives >= 2, 5e-10 / 1e-10 below:
    inside the north-star's 1e-9. The worst item is cross-checked with the mpmath solve itself."""
    from exact_solver import exact_solve

    B = len(pos)
    mask, v0 = po.canonical_mask_values(pos[0])
    values = np.zeros((B,) + v0.shape)
    values[:, :, 0, :] = pos
    ce, cost_e, _ = po.solve_exact128_batch(times, mask, values, N=N, derivative=der, n_threads=8)
    e_gpu = normwise(coeffs, ce).reshape(B, -1).max(axis=1)
    e_ora = normwise(ref_c, ce).reshape(B, -1).max(axis=1)
    # min-snap / jerk / acceleration: an order of magnitude inside the bar; velocity and position costs (cond(R_pp)
    # grows as the cost derivative drops): inside it by 2x
    assert e_gpu.max() < (1e-10 if der >= 2 else 5e-10), (int(e_gpu.argmax()), e_gpu.max(), e_ora.max())
    assert (np.abs(cost - cost_e) / np.abs(cost_e)).max() <= (1e-11 if der >= 2 else 1e-10)
    # where the CUDA path and the oracle visibly disagree it is the reference order's rounding noise
    err = normwise(coeffs, ref_c).reshape(B, -1).max(axis=1)
    vis = err > 1e-10
    assert np.all(e_ora[vis] > e_gpu[vis])
    for b in np.argsort(e_gpu)[-n:]:
        cm, cost_m, _ = exact_solve(N, der, times[b], mask, values[b])
        assert normwise(ce[b], cm).max() < 1e-15 and abs(cost_e[b] - cost_m) <= 1e-15 * abs(cost_m)


def gpu_solve(pos, times, derivative=4, end=None, device=True, want_free=True, layout="soa"):
    """pos [B,K+1,D], times [B,K], end [B,2,4,D] (AoS) -> coeffs [B,K,D,N], cost [B],
    free [B,D,K-1,4] (getFreeConstraints order), status [B]."""
    c = ctx()
    if layout == "soa":
        p, t = soa(pos), soa(times)
        e = soa(end) if end is not None else None
    else:
        p, t = np.ascontiguousarray(pos), np.ascontiguousarray(times)
        e = np.ascontiguousarray(end) if end is not None else None
    if device:
        p, t = dev(p), dev(t)
        e = dev(e) if e is not None else None
    r = c.solve_batch(p, t, e, N=N, derivative=derivative, want_free=want_free, layout=layout)
    if device:
        import torch

        torch.cuda.synchronize()
    conv = aos if layout == "soa" else (lambda x: x)
    return (conv(host(r["coeffs"])), host(r["cost"]), conv(host(r["free"])) if want_free else None,
            host(r["status"]))


@pytest.mark.parametrize("name", [n for n, p in REFERENCE_PARAMS.items()])
def test_reference_parameter_sets(po, name):
    """Every parameter set of TEST_OPT:754-851 (incl. K = 50, 75; D = 1, 3; snap/jerk/accel)."""
    prob = make_reference_problem(name)
    pos = prob["values"][:, 0, :][None]
    coeffs, cost, free, status = gpu_solve(pos, prob["times"][None], prob["derivative"])
    s = po.solve(N, prob["derivative"], prob["times"], prob["mask"], prob["values"])
    assert status[0] == 0
    # the reference order is noisier than 1e-9 for min-accel (see ORACLE_NOISE)
    tol = TOL if prob["derivative"] == 4 else ORACLE_NOISE[prob["derivative"]]
    assert normwise(coeffs[0], s.coeffs).max() < tol
    assert abs(cost[0] - s.cost) <= tol * abs(s.cost)
    if s.n_free:
        # getFreeConstraints order: vertex-major, derivative-minor (LIN_H:289-296)
        K, D = prob["K"], prob["D"]
        want = s.d_p.reshape(D, K - 1, 4)
        assert np.abs(free[0] - want).max() <= tol * np.abs(want).max()


def test_golden_fixtures_and_exact_solution():
    g = np.load(os.path.join(GOLD, "reference_params.npz"))
    for name in g["names"]:
        values = g[f"{name}/values"]
        times = g[f"{name}/times"]
        der = int(g[f"{name}/derivative"])
        coeffs, cost, _, status = gpu_solve(values[:, 0, :][None], times[None], der)
        assert status[0] == 0
        tol = TOL if der == 4 else ORACLE_NOISE[der]
        assert normwise(coeffs[0], g[f"{name}/oracle_coeffs"]).max() < tol, name
        assert abs(cost[0] - float(g[f"{name}/oracle_cost"])) <= tol * abs(cost[0]), name
        # against the 60-digit solution the closed-form path is much tighter
        assert normwise(coeffs[0], g[f"{name}/exact_coeffs"]).max() < 1e-10, name
        assert abs(cost[0] - float(g[f"{name}/exact_cost"])) <= 1e-11 * abs(cost[0]), name


def test_two_vertices_setup_golden():
    """TwoVerticesSetup TEST_OPT:707-751 through the CUDA path (K = 1, fully constrained)."""
    g = np.load(os.path.join(GOLD, "two_vertices_setup.npz"))
    pos = g["values"][:, 0, :][None]
    coeffs, cost, _, status = gpu_solve(pos, g["times"][None], 4, want_free=False)
    assert status[0] == 0
    assert np.abs(coeffs[0, 0, 0] - g["matlab_coeffs"]).max() < 5e-14
    exact = np.array([0, 0, 0, 0, 0, 0.2016, -0.1344, 0.03456, -0.004032, 0.0001792])
    assert np.abs(coeffs[0, 0, 0] - exact).max() < 1e-15


@pytest.mark.parametrize("layout", ["soa", "aos"])
@pytest.mark.parametrize("K,D,der", [(10, 3, 4), (10, 1, 4), (5, 3, 3), (5, 3, 2), (2, 3, 4), (3, 2, 4),
                                     (4, 3, 4), (7, 4, 4), (9, 3, 1), (6, 1, 0)])
def test_random_batch_vs_oracle(po, K, D, der, layout):
    B = 2048 if (K, D, der) == (10, 3, 4) else 256
    pos, times = random_problems(po, B, K, D, seed0=1000)
    coeffs, cost, free, status = gpu_solve(pos, times, der, layout=layout)
    ref_c, ref_cost = po.solve_canonical_batch(pos, times, N=N, derivative=der, n_threads=8)
    assert np.all(status == 0)
    tol = TOL if der == 4 else ORACLE_NOISE[der]
    err = normwise(coeffs, ref_c)
    assert err.max() < tol, (err.max(), np.unravel_index(err.argmax(), err.shape))
    assert (np.abs(cost - ref_cost) / np.abs(ref_cost)).max() < tol
    # the oracle tolerance above is the reference order's own noise (ORACLE_NOISE); the 1e-9 bar itself is
    # enforced on every item against the exact solution
    arbitrate(po, pos, times, der, coeffs, cost, ref_c, ref_cost)


def test_box_50_and_short_segments(po):
    """ConstraintPacking's +-50 m box (TEST_OPT:515-523) and segments down to ~0.6 s."""
    pos, times = random_problems(po, 256, 10, 3, box=50.0, seed0=12345)
    coeffs, cost, _, status = gpu_solve(pos, times, 4)
    ref_c, ref_cost = po.solve_canonical_batch(pos, times, n_threads=8)
    assert np.all(status == 0)
    assert normwise(coeffs, ref_c).max() < TOL
    pos, times = random_problems(po, 256, 10, 3, box=0.3, seed0=7)
    assert times.min() < 1.0
    coeffs, cost, _, status = gpu_solve(pos, times, 4)
    ref_c, ref_cost = po.solve_canonical_batch(pos, times, n_threads=8)
    assert np.all(status == 0)
    # sub-second segments: the reference order is ~3e-9 from exact here
    assert normwise(coeffs, ref_c).max() < ORACLE_NOISE[4]
    assert (np.abs(cost - ref_cost) / np.abs(ref_cost)).max() < TOL
    arbitrate(po, pos, times, 4, coeffs, cost, ref_c, ref_cost)


def test_nonzero_end_derivatives(po):
    """Start/end vertices with non-zero velocity..snap constraints (general makeStartOrEnd values)."""
    rng = np.random.RandomState(5)
    B, K, D = 64, 6, 3
    pos, times = random_problems(po, B, K, D, seed0=50)
    end = rng.uniform(-1, 1, size=(B, 2, 4, D))
    coeffs, cost, free, status = gpu_solve(pos, times, 4, end=end)
    assert np.all(status == 0)
    for b in range(B):
        mask, values = po.canonical_mask_values(pos[b])
        values[0, 1:, :] = end[b, 0]
        values[-1, 1:, :] = end[b, 1]
        s = po.solve(N, 4, times[b], mask, values)
        assert normwise(coeffs[b], s.coeffs).max() < TOL
        assert abs(cost[b] - s.cost) <= TOL * abs(s.cost)


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_host_memory_mode_matches_device_mode(po, layout):
    pos, times = random_problems(po, 777, 10, 3, seed0=3)
    a = gpu_solve(pos, times, 4, device=True, layout=layout)
    os.environ["MTG_HOST_CHUNK"] = "200"   # force several ragged chunks through all staging slots
    try:
        b = gpu_solve(pos, times, 4, device=False, layout=layout)
    finally:
        del os.environ["MTG_HOST_CHUNK"]
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    c = gpu_solve(pos, times, 4, device=True, layout="soa")
    for x, y in zip(a, c):   # the layout never changes a result bit
        assert np.array_equal(x, y)


def test_translation_invariance(po):
    """H annihilates constant offsets: a 1e6 m shift of all vertices changes only c_0."""
    pos, times = random_problems(po, 64, 10, 3, seed0=11)
    c0, cost0, free0, _ = gpu_solve(pos, times, 4)
    c1, cost1, free1, _ = gpu_solve(pos + 1.0e6, times, 4)
    # the shifted inputs differ from the originals by their own rounding (1e6 * 2^-53 ~ 1e-10 m)
    assert normwise(c1[..., 1:], c0[..., 1:]).max() < 1e-8
    assert np.allclose(cost0, cost1, rtol=1e-8)
    assert np.abs(free0 - free1).max() < 1e-8 * np.abs(free0).max()
    assert np.allclose(c1[..., 0] - 1.0e6, c0[..., 0], rtol=0, atol=1e-9)
    # an exactly representable shift leaves every derived quantity bit-identical
    c2, cost2, free2, _ = gpu_solve(np.round(pos * 1024) / 1024 + 4096.0, times, 4)
    c3, cost3, free3, _ = gpu_solve(np.round(pos * 1024) / 1024, times, 4)
    assert np.array_equal(c2[..., 1:], c3[..., 1:]) and np.array_equal(cost2, cost3)
    assert np.array_equal(free2, free3)


def test_status_flags_and_argument_errors(po):
    import mav_tube_trajectory_generation_b200 as m

    pos, times = random_problems(po, 8, 4, 3, seed0=9)
    times[3, 2] = 0.0
    times[5, 0] = -1.0
    coeffs, cost, _, status = gpu_solve(pos, times, 4)
    assert status[3] & 1 and status[5] & 1
    assert np.all(status[[0, 1, 2, 4, 6, 7]] == 0)
    with pytest.raises(m.MtgError):   # LIN_I:50-55
        gpu_solve(pos, times, 5)
    # empty batch is a no-op
    c = ctx()
    r = c.solve_batch(np.zeros((5, 3, 0)), np.zeros((4, 0)))
    assert r["coeffs"].shape == (4, 3, 10, 0)
    # inf / nan times
    times[3, 2] = np.inf
    times[5, 0] = np.nan
    _, _, _, status = gpu_solve(pos, times, 4)
    assert status[3] & 1 and status[5] & 1


def _powers(t, n):
    return t[..., None] ** np.arange(n)


def test_full_size_properties(po):
    """BASELINE config 2 size (65,536 x 10 segments x 3-D): size-independent properties
    of the reference's checkPath (TEST_OPT:113-195) and computeCost, vectorised in numpy."""
    import math

    B, K, D = 65536, 10, 3
    rng = np.random.RandomState(0)
    pos = rng.uniform(-10, 10, size=(B, K + 1, D))
    dist = np.linalg.norm(np.diff(pos, axis=1), axis=2)
    times = dist / 3.0 * 2 * (1.0 + 6.5 * 3.0 / 5.0 * np.exp(-dist / 3.0 * 2))
    coeffs, cost, free, status = gpu_solve(pos, times, 4)
    assert np.all(status == 0)
    Bt = po.base_coefficients()
    # derivative k of every polynomial at t = 0 and t = T
    def deriv_at(tvals, k):
        # coeffs [B,K,D,N]; tvals [B,K]
        j = np.arange(k, N)
        w = Bt[k, k:N] * tvals[..., None] ** (j - k)            # [B,K,N-k]
        return np.einsum("bkdn,bkn->bkd", coeffs[..., k:], w)
    zero = np.zeros_like(times)
    scale = np.abs(pos).max()
    for k in range(5):
        at0 = deriv_at(zero, k)
        atT = deriv_at(times, k)
        if k == 0:
            assert np.abs(at0 - pos[:, :-1]).max() < 1e-6
            assert np.abs(atT - pos[:, 1:]).max() < 1e-6
        else:
            assert np.abs(at0[:, 0]).max() < 1e-6 and np.abs(atT[:, -1]).max() < 1e-6
        assert np.abs(atT[:, :-1] - at0[:, 1:]).max() < 1e-6 * max(1.0, scale)
        if k >= 1:  # and the solved free derivatives are those values
            want = np.moveaxis(free[:, :, :, k - 1], 1, 2)   # [B,D,K-1] -> [B,K-1,D]
            assert np.abs(at0[:, 1:] - want).max() < 1e-9 * max(1.0, np.abs(free).max())
    # cost == 0.5 sum c^T Q c with Q of LIN_I:557-573
    d = 4
    a = np.arange(d, N)
    expo = a[:, None] + a[None, :] - 2 * d + 1
    q = (Bt[d, d:N][:, None] * Bt[d, d:N][None, :]) * 2.0 / expo
    Q = q[None, None] * times[..., None, None] ** expo[None, None]
    c4 = coeffs[..., d:]
    want = 0.5 * np.einsum("bkdi,bkij,bkdj->b", c4, Q, c4)
    assert (np.abs(want - cost) / cost).max() < 1e-8
    # 4,096 of them against the oracle (1e-9) and against the exact solution (1e-10)
    idx = np.sort(rng.choice(B, 4096, replace=False))
    ref_c, ref_cost = po.solve_canonical_batch(pos[idx], times[idx], n_threads=8)
    assert normwise(coeffs[idx], ref_c).max() < TOL
    assert (np.abs(cost[idx] - ref_cost) / ref_cost).max() < TOL
    arbitrate(po, pos[idx], times[idx], 4, coeffs[idx], cost[idx], ref_c, ref_cost)
    assert math.isfinite(cost.sum())


def test_launch_counter_moves():
    c = ctx()
    n0 = c.launch_count
    pos = np.random.RandomState(1).uniform(-5, 5, size=(4, 3, 3))
    gpu_solve(pos, np.full((4, 2), 2.0), 4)
    assert c.launch_count == n0 + 1


def test_cost_only_solve(po):
    """coeffs = NULL: the sweep form of the solve (cost, status and d_p only); same numbers."""
    pos, times = random_problems(po, 1000, 10, 3, seed0=888)
    c = ctx()
    p, t = dev(soa(pos)), dev(soa(times))
    full = c.solve_batch(p, t, want_free=True)
    lean = c.solve_batch(p, t, want_free=True, want_coeffs=False)
    assert lean["coeffs"] is None
    assert np.array_equal(host(lean["cost"]), host(full["cost"]))
    assert np.array_equal(host(lean["free"]), host(full["free"]))
    assert np.array_equal(host(lean["status"]), host(full["status"]))
    hl = c.solve_batch(soa(pos), soa(times), want_coeffs=False)       # host-memory mode
    assert np.array_equal(hl["cost"], host(full["cost"]))
    import mav_tube_trajectory_generation_b200 as m

    with pytest.raises(m.MtgError):
        c.solve_batch(p, t, want_coeffs=False, want_cost=False, want_free=False)

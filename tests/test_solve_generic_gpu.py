"""-m gpu parity tests of mtg_solve_generic_batch: arbitrary (batch-shared) constraint patterns —
the general setupConstraintReorderingMatrix / solveLinear (LIN_I:171-252, 337-379) — against the
oracle (which handles any mask) and the 60-digit solve. Bars as in test_solve_gpu.py."""
import numpy as np
import pytest

from gpu_util import aos, ctx, dev, host, normwise, random_problems, soa

pytestmark = pytest.mark.gpu
N, H = 10, 5


def gpu_generic(mask, values, times, derivative=4, layout="soa", device=True):
    """values [B,K+1,H,D], times [B,K] -> coeffs [B,K,D,N], cost [B], free [B,D,n_free], status."""
    c = ctx()
    conv_in = soa if layout == "soa" else np.ascontiguousarray
    v, t = conv_in(values), conv_in(times)
    if device:
        v, t = dev(v), dev(t)
    r = c.solve_generic_batch(mask, v, t, N=N, derivative=derivative, layout=layout)
    if device:
        import torch

        torch.cuda.synchronize()
    conv = aos if layout == "soa" else (lambda x: x)
    free = conv(host(r["free"])) if r["free"] is not None else None
    return conv(host(r["coeffs"])), host(r["cost"]), free, host(r["status"])


def make_values(po, B, K, D, mask, seed):
    """Random vertex positions (createRandomVertices recipe) + modest random values for every other
    fixed derivative."""
    pos, times = random_problems(po, B, K, D, seed0=seed)
    rng = np.random.RandomState(seed)
    values = rng.uniform(-1.0, 1.0, size=(B, K + 1, H, D)) * np.array([1.0, 1.0, 0.5, 0.25, 0.1])[None, None, :, None]
    values[:, :, 0, :] = pos
    return values * mask[None, :, :, None], times


def canonical_mask(K):
    m = np.zeros((K + 1, H), dtype=np.uint8)
    m[:, 0] = 1
    m[0, :] = 1
    m[K, :] = 1
    return m


PATTERNS = {
    # interior velocity fixed at every second vertex
    "interior_velocity": lambda K: _with(canonical_mask(K), [(v, 1) for v in range(2, K, 2)]),
    # goal: only position fixed (free final velocity .. snap)
    "free_goal_derivatives": lambda K: _without(canonical_mask(K), [(K, k) for k in range(1, H)]),
    # start: position + velocity only; one interior vertex fixes position, velocity and acceleration
    "mixed": lambda K: _with(_without(canonical_mask(K), [(0, 2), (0, 3), (0, 4)]), [(K // 2, 1), (K // 2, 2)]),
    # an interior vertex that does not even fix its position (a pure "knot")
    "free_knot": lambda K: _without(canonical_mask(K), [(K // 2, 0)]),
    # everything fixed at one interior vertex: the chain decouples there (f_v = 0)
    "fully_fixed_interior": lambda K: _with(canonical_mask(K), [(K // 2, k) for k in range(1, H)]),
}


def _with(m, entries):
    for v, k in entries:
        m[v, k] = 1
    return m


def _without(m, entries):
    for v, k in entries:
        m[v, k] = 0
    return m


@pytest.mark.parametrize("pattern", list(PATTERNS))
@pytest.mark.parametrize("K,D,der", [(10, 3, 4), (5, 1, 3), (7, 2, 2)])
def test_patterns_vs_oracle_and_exact(po, pattern, K, D, der):
    from exact_solver import exact_solve

    mask = PATTERNS[pattern](K)
    B = 24
    values, times = make_values(po, B, K, D, mask, seed=hash(pattern) % 1000 + K)
    coeffs, cost, free, status = gpu_generic(mask, values, times, der)
    assert np.all(status == 0)
    errs = np.zeros(B)
    for b in range(B):
        s = po.solve(N, der, times[b], mask, values[b], want_R=False)
        errs[b] = normwise(coeffs[b], s.coeffs).max()
        assert errs[b] < 1e-6, (b, errs[b])              # the oracle's (= the reference's) own order is this noisy
        assert abs(cost[b] - s.cost) <= 1e-6 * s.cost
        assert np.abs(free[b] - s.d_p.reshape(D, -1)).max() <= 1e-6 * max(1.0, np.abs(s.d_p).max())
    # the 1e-9 bar on EVERY item, against the exact (binary128) solution of the same normal equations
    ce, cost_e, dpe = po.solve_exact128_batch(times, mask, values, N=N, derivative=der, n_threads=8)
    assert normwise(coeffs, ce).max() < 1e-9, normwise(coeffs, ce).max()
    assert (np.abs(cost - cost_e) / np.abs(cost_e)).max() <= 1e-9
    assert np.abs(free - dpe).max() <= 1e-9 * max(1.0, np.abs(dpe).max())
    b = int(np.argmax(errs))                               # and the arbiter itself against the 60-digit solve
    cm, cost_m, _ = exact_solve(N, der, times[b], mask, values[b])
    assert normwise(ce[b], cm).max() < 1e-15 and abs(cost_e[b] - cost_m) <= 1e-15 * abs(cost_m)


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_canonical_pattern_equals_solve_batch(po, layout):
    K, D, B = 10, 3, 512
    mask = canonical_mask(K)
    pos, times = random_problems(po, B, K, D, seed0=77)
    values = np.zeros((B, K + 1, H, D))
    values[:, :, 0, :] = pos
    coeffs, cost, free, status = gpu_generic(mask, values, times, 4, layout=layout)
    c = ctx()
    r = c.solve_batch(dev(soa(pos)), dev(soa(times)), want_free=True)
    ref_c, ref_cost, ref_free = aos(host(r["coeffs"])), host(r["cost"]), aos(host(r["free"]))
    assert np.all(status == 0)
    assert normwise(coeffs, ref_c).max() < 1e-10
    assert np.allclose(cost, ref_cost, rtol=1e-10, atol=0)
    assert np.abs(free - ref_free.reshape(B, D, -1)).max() <= 1e-9 * np.abs(ref_free).max()
    # host-memory mode == device mode
    ch, costh, freeh, sth = gpu_generic(mask, values[:100], times[:100], 4, layout=layout, device=False)
    assert np.array_equal(ch, coeffs[:100]) and np.array_equal(costh, cost[:100])


def test_two_vertices_fully_constrained_and_underdetermined(po):
    # TwoVerticesSetup (TEST_OPT:707-751): n_free = 0
    mask = np.ones((2, H), dtype=np.uint8)
    values = np.zeros((1, 2, H, 1))
    values[0, 1, 0, 0] = 5.0
    coeffs, cost, free, status = gpu_generic(mask, values, np.array([[5.0]]), 4)
    gold = [0.2016, -0.1344, 0.03456, -0.004032, 0.0001792]
    assert status[0] == 0 and free is None
    assert np.abs(coeffs[0, 0, 0, 5:] - gold).max() < 1e-12 and np.abs(coeffs[0, 0, 0, :5]).max() < 1e-12
    # no position fixed anywhere: R_pp is singular (translations). A pivot that comes out non-positive
    # (or below 1e-11 of its pre-elimination diagonal) raises the status bit; rank deficiency hidden by
    # rounding cannot be detected by any factorisation (the reference's SparseQR returns a basic
    # solution silently). Never an abort, never a NaN.
    K = 4
    mask = canonical_mask(K)
    mask[:, 0] = 0
    values, times = make_values(po, 4, K, 3, mask, seed=5)
    coeffs, _, _, status = gpu_generic(mask, values, times, 4)
    assert np.any(status & 2) and np.all(np.isfinite(coeffs))


def test_coeffs_from_derivatives_round_trip(po):
    """C [d_f; d_p] -> coefficients (LIN_I:254-275) for a general pattern: merging the solved free
    derivatives back into the fixed ones reproduces the solve's coefficients and cost."""
    K, D, B = 8, 3, 64
    mask = PATTERNS["mixed"](K)
    values, times = make_values(po, B, K, D, mask, seed=9)
    coeffs, cost, free, status = gpu_generic(mask, values, times, 4)
    full = values.copy()
    q = 0
    for v in range(K + 1):
        for k in range(H):
            if not mask[v, k]:
                full[:, v, k, :] = free[:, :, q]
                q += 1
    c = ctx()
    r = c.coeffs_from_derivatives_batch(dev(soa(full)), dev(soa(times)))
    assert np.array_equal(aos(host(r["coeffs"])), coeffs)
    assert np.allclose(host(r["cost"]), cost, rtol=1e-14, atol=0)
    ra = c.coeffs_from_derivatives_batch(np.ascontiguousarray(full), np.ascontiguousarray(times), layout="aos")
    assert np.array_equal(ra["coeffs"], coeffs)

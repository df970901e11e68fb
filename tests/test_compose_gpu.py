"""-m gpu parity tests of the batched trajectory composition / I/O entry points (N3) against the oracle:
getVertexAtTime (TRAJ_C:248-262), dimension split / append (TRAJ_C:136-182), addTrajectories (TRAJ_C:230-246),
computeCost of given coefficients (LIN_I:113-130) and the sample dump of printMatlabSampledTrajectory
(NL_I:2907-3003)."""
import numpy as np
import pytest

from gpu_util import aos, ctx, dev, host, random_problems, soa

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def solved(po):
    pos, times = random_problems(po, 96, 10, 3, seed0=7300)
    coeffs, cost = po.solve_canonical_batch(pos, times, n_threads=8)
    return pos, times, coeffs, cost


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_vertex_at_time(po, solved, layout):
    _, times, coeffs, _ = solved
    B, K = times.shape
    conv_in = soa if layout == "soa" else np.ascontiguousarray
    conv = aos if layout == "soa" else (lambda x: x)
    tmax = np.array([float(host(ctx().max_time_batch(dev(soa(times))))[b]) for b in range(B)])
    rng = np.random.RandomState(1)
    t = rng.uniform(0, 1, size=B) * tmax
    t[0], t[1], t[2] = 0.0, tmax[1], tmax[2] * 1.5          # start vertex, goal vertex, out of range
    t[3] = times[3, 0] + times[3, 1]                         # exactly on a vertex: the right-hand segment
    c = ctx()
    for device in (True, False):
        f = dev if device else (lambda x: x)
        r = c.vertex_at_time_batch(f(conv_in(coeffs)), f(conv_in(times)), f(t), 4, layout=layout)
        out, seg, st = conv(host(r["out"])), host(r["segment_idx"]), host(r["status"])
        for b in range(B):
            if b == 2:
                assert st[b] == 4 and seg[b] == -1 and np.all(out[b] == 0.0)
                continue
            want = po.vertex_at_time(coeffs[b], times[b], t[b], 4)
            _, s = po.traj_evaluate(coeffs[b], times[b], t[b], 0)
            assert seg[b] == s and st[b] == 0
            # derivative k at tau is a sum of terms B(k, j) |c_j| tau^(j-k): compare against that size (at a
            # rest-to-rest end the derivatives themselves are ~0)
            i, tau = s, t[b] - times[b, :s].sum()
            j = np.arange(10)
            for k in range(5):
                size = sum(np.abs(coeffs[b, i, :, jj]) * np.prod(np.arange(jj - k + 1, jj + 1)) * abs(tau) ** (jj - k)
                           for jj in range(k, 10)) + 1e-300
                assert np.all(np.abs(out[b, k] - want[k]) <= 1e-9 * size)
        assert seg[1] == K - 1 and seg[3] == 2


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_dimension_split_append_and_concat(po, solved, layout):
    _, times, coeffs, _ = solved
    conv_in = soa if layout == "soa" else np.ascontiguousarray
    conv = aos if layout == "soa" else (lambda x: x)
    c = ctx()
    a = dev(conv_in(coeffs))
    # getTrajectoryWithSingleDimension(1)
    one = c.pick_dimensions_batch(a, [1], layout=layout)
    assert np.array_equal(conv(host(one)), coeffs[:, :, 1:2, :])
    # getTrajectoryWithAppendedDimension: [x, y, z] + [y] -> 4-D
    four = c.pick_dimensions_batch(a, [0, 1, 2, 3], coeffs_b=one, layout=layout)
    assert np.array_equal(conv(host(four)), np.concatenate([coeffs, coeffs[:, :, 1:2, :]], axis=2))
    # a permutation through host memory
    perm = c.pick_dimensions_batch(conv_in(coeffs), [2, 0], layout=layout)
    assert np.array_equal(conv(perm), coeffs[:, :, [2, 0], :])
    # addTrajectories: segments 0..3 + 4..9 + 0..1
    parts = [coeffs[:, :4], coeffs[:, 4:], coeffs[:, :2]]
    tparts = [times[:, :4], times[:, 4:], times[:, :2]]
    for device in (True, False):
        f = dev if device else (lambda x: x)
        oc, ot = c.concat_segments_batch([f(conv_in(x)) for x in parts], [f(conv_in(x)) for x in tparts], layout=layout)
        assert np.array_equal(conv(host(oc)), np.concatenate(parts, axis=1))
        assert np.array_equal(conv(host(ot)), np.concatenate(tparts, axis=1))
    import mav_tube_trajectory_generation_b200 as m

    with pytest.raises(m.MtgError):       # CHECK_LT(dimension, D_), TRAJ_C:137
        c.pick_dimensions_batch(a, [3], layout=layout)


def test_compute_cost_of_given_coefficients(po, solved):
    """computeCost (LIN_I:113-130) on the solve's coefficients equals the solve's cost; after a change of the
    segment times WITHOUT a new solve it is the reference's 0.5 sum c^T Q(T_new) c."""
    pos, times, coeffs, cost = solved
    c = ctx()
    r = c.compute_cost_batch(dev(soa(coeffs)), dev(soa(times)))
    assert np.allclose(host(r["cost"]), cost, rtol=1e-9, atol=0)
    t2 = times * 1.25
    r2 = c.compute_cost_batch(np.ascontiguousarray(coeffs), np.ascontiguousarray(t2), layout="aos")
    for b in range(0, len(cost), 7):
        want = 0.0
        for i in range(times.shape[1]):
            Q = po.quadratic_cost_jacobian(10, 4, t2[b, i])
            for d_ in range(3):
                want += 0.5 * coeffs[b, i, d_] @ Q @ coeffs[b, i, d_]
        assert abs(r2["cost"][b] - want) <= 1e-9 * want
    for der in (2, 3):
        r3 = c.compute_cost_batch(dev(soa(coeffs[:8])), dev(soa(times[:8])), derivative=der)
        for b in range(8):
            want = sum(0.5 * coeffs[b, i, d_] @ po.quadratic_cost_jacobian(10, der, times[b, i]) @ coeffs[b, i, d_]
                       for i in range(10) for d_ in range(3))
            assert abs(host(r3["cost"])[b] - want) <= 1e-9 * want


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_sample_dump(po, solved, layout):
    """printMatlabSampledTrajectory (NL_I:2907-3003): same row count (the t += dt accumulation is replayed),
    same times bit for bit, values to 1e-9 of each column's scale, tm column and zero rows like the reference."""
    _, times, coeffs, _ = solved
    B = 24
    conv_in = soa if layout == "soa" else np.ascontiguousarray
    dt = 0.01
    max_rows = int(sum(np.ceil(times[:B].max(axis=0) / dt) + 1)) + 4
    c = ctx()
    r = c.sample_dump_batch(dev(conv_in(coeffs[:B])), dev(conv_in(times[:B])), dt, max_rows, layout=layout)
    rows, n = host(r["rows"]), host(r["n_rows"])
    assert np.all(host(r["status"]) == 0)
    for b in range(B):
        want, k = po.sample_dump(coeffs[b], times[b], dt, max_rows)
        assert n[b] == k
        assert np.array_equal(rows[b, :, 0], want[:, 0])                       # sample times
        assert np.array_equal(rows[b, :, -1], want[:, -1])                     # tm column (row i = end of segment i)
        assert np.all(rows[b, k:, :-1] == 0.0)
        scale = np.abs(want[:k]).max(axis=0) + 1e-300
        assert (np.abs(rows[b, :k] - want[:k]) / scale).max() <= 1e-9
    # too few rows: truncated like the reference's `if (j < output.rows())`, flagged
    r = c.sample_dump_batch(dev(conv_in(coeffs[:2])), dev(conv_in(times[:2])), dt, 100, layout=layout)
    assert np.all(host(r["n_rows"]) == 100) and np.all(host(r["status"]) == 8)
    # host-memory mode
    rh = c.sample_dump_batch(conv_in(coeffs[:B]), conv_in(times[:B]), dt, max_rows, layout=layout)
    assert np.array_equal(rh["rows"], rows) and np.array_equal(rh["n_rows"], n)

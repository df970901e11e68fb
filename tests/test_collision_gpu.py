"""-m gpu parity tests of mtg_collision_cost_batch (N4: the collision potential of the non-linear layer,
NL_I:1608-1780 / 1783-1917 / 2659-2684, against a DENSE distance grid instead of the supereight octree) against
the oracle's restatement of the same loops on the same grid."""
import numpy as np
import pytest

from gpu_util import aos, ctx, dev, host, random_problems, soa

pytestmark = pytest.mark.gpu
N, NF = 10, 4


def sphere_field(res, lo, hi, centres, radii):
    """Distance (m) to the nearest of a few spherical obstacles, sampled at voxel indices (like the reference's
    distance to the nearest occupied voxel times the resolution). Returns (grid [nx,ny,nz], origin voxel)."""
    o = np.floor(np.asarray(lo) / res).astype(int) - 2
    n = (np.ceil(np.asarray(hi) / res).astype(int) + 2) - o
    ix, iy, iz = np.meshgrid(*(np.arange(k) for k in n), indexing="ij")
    p = (np.stack([ix, iy, iz], axis=-1) + o) * res
    d = np.full(tuple(n), np.inf)
    for c, r in zip(centres, radii):
        d = np.minimum(d, np.maximum(np.linalg.norm(p - np.asarray(c), axis=-1) - r, 0.0))
    return np.ascontiguousarray(d), o


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_collision_cost_and_gradient_vs_oracle(po, layout):
    B, K = 96, 6
    pos, times = random_problems(po, B, K, 3, box=5.0, seed0=6600)
    coeffs, _ = po.solve_canonical_batch(pos, times, n_threads=8)
    res, dt = 0.2, 0.05
    lo, hi = [-8.0] * 3, [8.0] * 3
    rng = np.random.RandomState(4)
    centres = rng.uniform(-5, 5, size=(12, 3))
    radii = rng.uniform(0.1, 0.6, size=12)
    grid, origin = sphere_field(res, lo, hi, centres, radii)
    conv_in = soa if layout == "soa" else np.ascontiguousarray
    conv = aos if layout == "soa" else (lambda x: x)
    c = ctx()
    kw = dict(epsilon=1.5, robot_radius=0.3, multiplier=2.0)
    r = c.collision_cost_batch(dev(conv_in(coeffs)), dev(conv_in(times)), dev(grid), origin, res, lo, hi, dt, layout=layout,
                               **kw)
    J, g, col, chk = host(r["J_c"]), conv(host(r["grad"])), host(r["in_collision"]), host(r["n_checks"])
    assert np.all(host(r["status"]) == 0)
    n_free_cost = 0
    for b in range(B):
        Jo, go, co, ko = po.collision_cost(coeffs[b], times[b], grid, origin, res, lo, hi, dt, **kw)
        assert bool(col[b]) == co and chk[b] == ko, (b, col[b], co, chk[b], ko)
        assert abs(J[b] - Jo) <= 1e-9 * max(Jo, 1e-12)
        scale = np.abs(go).max()
        assert np.abs(g[b].reshape(3, -1) - go).max() <= 1e-9 * scale + 1e-15
        n_free_cost += int((not co) and Jo > 0)
    assert col.sum() > 3 and (col == 0).sum() > 3 and n_free_cost > 3       # the batch exercises every branch
    # cost only
    r2 = c.collision_cost_batch(dev(conv_in(coeffs)), dev(conv_in(times)), dev(grid), origin, res, lo, hi, dt,
                                layout=layout, want_grad=False, **kw)
    assert np.array_equal(host(r2["J_c"]), J)


def test_bounds_and_free_space(po):
    """No obstacle in reach: zero cost, zero gradient; a trajectory that leaves [min_bound + res, max_bound - res]
    is in collision (NL_I:1800-1807)."""
    B, K = 16, 4
    pos, times = random_problems(po, B, K, 3, box=3.0, seed0=6700)
    coeffs, _ = po.solve_canonical_batch(pos, times, n_threads=8)
    grid = np.full((8, 8, 8), np.inf)
    c = ctx()
    r = c.collision_cost_batch(dev(soa(coeffs)), dev(soa(times)), dev(grid), [0, 0, 0], 0.1, [-50] * 3, [50] * 3, 0.05)
    assert np.all(host(r["J_c"]) == 0.0) and np.all(host(r["grad"]) == 0.0) and not host(r["in_collision"]).any()
    r = c.collision_cost_batch(dev(soa(coeffs)), dev(soa(times)), dev(grid), [0, 0, 0], 0.1, [-1.0] * 3, [1.0] * 3, 0.05)
    for b in range(B):
        Jo, _, co, ko = po.collision_cost(coeffs[b], times[b], grid, [0, 0, 0], 0.1, [-1.0] * 3, [1.0] * 3, 0.05)
        assert bool(host(r["in_collision"])[b]) == co and host(r["n_checks"])[b] == ko and host(r["J_c"])[b] == Jo
    assert host(r["in_collision"]).sum() > B // 2


@pytest.mark.parametrize("n_coeffs", [6, 12])
def test_other_polynomial_orders(po, n_coeffs):
    """The same comparison for N = 6 (2 free derivatives per vertex) and N = 12 (5; 36 moment sums per segment, more
    than one per lane), AoS records."""
    B, K = 32, 5
    pos, times = random_problems(po, B, K, 3, box=5.0, seed0=6800)
    coeffs, _ = po.solve_canonical_batch(pos, times, N=n_coeffs, derivative=n_coeffs // 2 - 1, n_threads=8)
    res, dt = 0.25, 0.07
    lo, hi = [-9.0] * 3, [9.0] * 3
    rng = np.random.RandomState(11)
    grid, origin = sphere_field(res, lo, hi, rng.uniform(-5, 5, size=(8, 3)), rng.uniform(0.1, 0.5, size=8))
    kw = dict(epsilon=2.0, robot_radius=0.2, multiplier=1.5)
    r = ctx().collision_cost_batch(dev(coeffs), dev(times), dev(grid), origin, res, lo, hi, dt, layout="aos", **kw)
    J, g, col, chk = host(r["J_c"]), host(r["grad"]), host(r["in_collision"]), host(r["n_checks"])
    assert np.all(host(r["status"]) == 0)
    for b in range(B):
        Jo, go, co, ko = po.collision_cost(coeffs[b], times[b], grid, origin, res, lo, hi, dt, **kw)
        assert bool(col[b]) == co and chk[b] == ko
        assert abs(J[b] - Jo) <= 1e-9 * max(Jo, 1e-12)
        assert np.abs(g[b].reshape(3, -1) - go).max() <= 1e-9 * np.abs(go).max() + 1e-15
    assert (col == 0).sum() > 3

"""TEST INFRASTRUCTURE — ctypes front-end of the CPU oracle (oracle/liboracle.so).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import
this module. It is the checker of the CUDA path, never part of it. See
``oracle/oracle.h`` for what each function restates (reference file:line) and
for which rows are pinned / unpinned.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)


def build(force: bool = False) -> str:
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.mtgo_poly_evaluate.restype = C.c_double
        _lib.mtgo_sampled_maximum_magnitude.restype = C.c_double
        _lib.mtgo_cost_numeric.restype = C.c_double
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def has_reference_rpoly() -> bool:
    return bool(lib().mtgo_has_reference_rpoly())


# ----------------------------------------------------------------- matrices
def base_coefficients() -> np.ndarray:
    out = np.empty((22, 22))
    lib().mtgo_base_coefficients(_d(out))
    return out


def quadratic_cost_jacobian(N: int, derivative: int, t: float) -> np.ndarray:
    Q = np.empty((N, N))
    lib().mtgo_quadratic_cost_jacobian(N, derivative, C.c_double(t), _d(Q))
    return Q


def setup_mapping_matrix(N: int, t: float) -> np.ndarray:
    A = np.empty((N, N))
    lib().mtgo_setup_mapping_matrix(N, C.c_double(t), _d(A))
    return A


def invert_mapping_matrix(A: np.ndarray) -> np.ndarray:
    A = _f64(A)
    out = np.empty_like(A)
    lib().mtgo_invert_mapping_matrix(A.shape[0], _d(A), _d(out))
    return out


def general_inverse(M: np.ndarray) -> np.ndarray:
    M = _f64(M)
    out = np.empty_like(M)
    rc = lib().mtgo_general_inverse(M.shape[0], _d(M), _d(out))
    if rc:
        raise ValueError("singular")
    return out


# --------------------------------------------------------------- generators
def create_random_vertices(maximum_derivative: int, n_segments: int, pos_min, pos_max,
                           seed: int, N: int = 10):
    """VTX_C:27-82. Returns (mask[(K+1),N/2] uint8, values[(K+1),N/2,D])."""
    pos_min = _f64(np.atleast_1d(pos_min))
    pos_max = _f64(np.atleast_1d(pos_max))
    D = pos_min.size
    h = N // 2
    mask = np.zeros((n_segments + 1, h), dtype=np.uint8)
    values = np.zeros((n_segments + 1, h, D))
    rc = lib().mtgo_create_random_vertices(maximum_derivative, n_segments, D, _d(pos_min),
                                           _d(pos_max), C.c_uint64(seed), h,
                                           mask.ctypes.data_as(_u8p), _d(values))
    if rc < 0:
        raise ValueError("createRandomVertices: invalid arguments")
    return mask, values


def estimate_segment_times_nfabian(positions, v_max, a_max, magic=6.5) -> np.ndarray:
    positions = _f64(positions)
    K = positions.shape[0] - 1
    D = positions.shape[1]
    out = np.empty(K)
    lib().mtgo_estimate_segment_times_nfabian(K, D, _d(positions), C.c_double(v_max),
                                              C.c_double(a_max), C.c_double(magic), _d(out))
    return out


def estimate_segment_times_velocity_ramp(positions, v_max, a_max) -> np.ndarray:
    positions = _f64(positions)
    K = positions.shape[0] - 1
    D = positions.shape[1]
    out = np.empty(K)
    lib().mtgo_estimate_segment_times_velocity_ramp(K, D, _d(positions), C.c_double(v_max),
                                                    C.c_double(a_max), _d(out))
    return out


# -------------------------------------------------------------------- solve
class Solution:
    __slots__ = ("coeffs", "cost", "n_all", "n_fixed", "n_free", "d_f", "d_p", "R", "col_of_row")


def solve(N, derivative, times, mask, values, solver: int = 0, want_R: bool = True) -> Solution:
    """P1..P8: mask[(K+1),N/2], values[(K+1),N/2,D], times[K] -> Solution."""
    times = _f64(times)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    values = _f64(values)
    K = times.size
    h = N // 2
    D = values.shape[2]
    assert mask.shape == (K + 1, h) and values.shape == (K + 1, h, D)
    n_all = K * N
    coeffs = np.zeros((K, D, N))
    cost = C.c_double(0.0)
    counts = (C.c_int * 3)()
    d_f = np.zeros((D, n_all))
    d_p = np.zeros((D, n_all))
    R = np.zeros((n_all, n_all)) if want_R else None
    col = np.zeros(n_all, dtype=np.int32)
    rc = lib().mtgo_solve(N, D, K, derivative, _d(times), mask.ctypes.data_as(_u8p), _d(values),
                          solver, _d(coeffs), C.byref(cost), counts, _d(d_f), _d(d_p),
                          _d(R) if want_R else None, col.ctypes.data_as(_ip))
    if rc:
        raise ValueError(f"oracle solve failed rc={rc}")
    s = Solution()
    s.coeffs, s.cost = coeffs, cost.value
    s.n_all, s.n_fixed, s.n_free = counts[0], counts[1], counts[2]
    s.d_f = d_f.reshape(-1)[: D * s.n_fixed].reshape(D, s.n_fixed).copy()
    s.d_p = d_p.reshape(-1)[: D * s.n_free].reshape(D, s.n_free).copy()
    n = s.n_fixed + s.n_free
    s.R = R.reshape(-1)[: n * n].reshape(n, n).copy() if want_R else None
    s.col_of_row = col[: s.n_all].copy()
    return s


def solve_canonical_batch(positions, times, N=10, derivative=4, solver=0, n_threads=1):
    """positions [B,K+1,D], times [B,K] -> coeffs [B,K,D,N], cost [B]."""
    positions = _f64(positions)
    times = _f64(times)
    B, Kp1, D = positions.shape
    K = Kp1 - 1
    coeffs = np.empty((B, K, D, N))
    cost = np.empty(B)
    rc = lib().mtgo_solve_canonical_batch(B, N, D, K, derivative, _d(positions), _d(times),
                                          solver, n_threads, _d(coeffs), _d(cost))
    if rc:
        raise ValueError("oracle batch solve failed")
    return coeffs, cost


def solve_exact128_batch(times, mask, values, N=10, derivative=4, n_threads=8):
    """Binary128 arbiter (oracle/exact128.cpp): the reference's normal equations solved in IEEE binary128 and
    rounded once. times [B,K], mask [K+1,h] shared, values [B,K+1,h,D] -> coeffs [B,K,D,N], cost [B], d_p [B,D,n_free]."""
    times = _f64(times)
    values = _f64(values)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    B, K = times.shape
    D = values.shape[-1]
    n_free = int((mask == 0).sum())
    coeffs = np.zeros((B, K, D, N))
    cost = np.zeros(B)
    d_p = np.zeros((B, D, max(n_free, 1)))
    bad = lib().mtgo_solve_exact128_batch(B, N, D, K, derivative, _d(times), mask.ctypes.data_as(_u8p), _d(values),
                                          _d(coeffs), _d(cost), _d(d_p) if n_free else None, n_free, n_threads)
    if bad:
        raise RuntimeError(f"{bad} problems could not be solved in binary128")
    return coeffs, cost, d_p[:, :, :n_free]


def canonical_mask_values(positions, N=10):
    """mask/values of the createRandomVertices pattern for given positions [K+1,D]."""
    positions = _f64(positions)
    Kp1, D = positions.shape
    h = N // 2
    mask = np.zeros((Kp1, h), dtype=np.uint8)
    values = np.zeros((Kp1, h, D))
    mask[:, 0] = 1
    values[:, 0, :] = positions
    mask[0, :] = 1
    mask[-1, :] = 1
    return mask, values


def coeffs_from_free_constraints(N, times, mask, values, d_p) -> np.ndarray:
    times = _f64(times)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    values = _f64(values)
    d_p = _f64(d_p)
    K = times.size
    D = values.shape[2]
    coeffs = np.zeros((K, D, N))
    rc = lib().mtgo_coeffs_from_free_constraints(N, D, K, _d(times), mask.ctypes.data_as(_u8p),
                                                 _d(values), _d(d_p), _d(coeffs))
    if rc:
        raise ValueError(rc)
    return coeffs


def cost_time_fd(N, derivative, times, mask, values, d_p, increment_time, central: bool):
    """P9. Returns (J_nominal, J_plus[K], J_minus[K] or None, grad_d[K])."""
    times = _f64(times)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    values = _f64(values)
    d_p = _f64(d_p)
    K = times.size
    D = values.shape[2]
    J0 = C.c_double(0.0)
    Jp = np.zeros(K)
    Jm = np.zeros(K)
    g = np.zeros(K)
    rc = lib().mtgo_cost_time_fd(N, D, K, derivative, _d(times), mask.ctypes.data_as(_u8p),
                                 _d(values), _d(d_p), C.c_double(increment_time),
                                 1 if central else 0, C.byref(J0), _d(Jp), _d(Jm), _d(g))
    if rc:
        raise ValueError(rc)
    return J0.value, Jp, (Jm if central else None), g


# --------------------------------------------------------------- evaluation
def poly_evaluate(c, t: float, derivative: int) -> float:
    c = _f64(c)
    return lib().mtgo_poly_evaluate(c.size, _d(c), C.c_double(t), derivative)


def poly_derivative_coefficients(c, derivative: int) -> np.ndarray:
    c = _f64(c)
    out = np.empty_like(c)
    lib().mtgo_poly_derivative_coefficients(c.size, _d(c), derivative, _d(out))
    return out


def convolve(data, kernel) -> np.ndarray:
    data = _f64(data)
    kernel = _f64(kernel)
    out = np.empty(data.size + kernel.size - 1)
    lib().mtgo_convolve(_d(data), data.size, _d(kernel), kernel.size, _d(out))
    return out


def traj_evaluate(coeffs, times, t: float, derivative: int):
    coeffs = _f64(coeffs)
    times = _f64(times)
    K, D, N = coeffs.shape
    out = np.zeros(D)
    seg = lib().mtgo_traj_evaluate(N, D, K, _d(coeffs), _d(times), C.c_double(t), derivative, _d(out))
    return out, seg


def traj_evaluate_range(coeffs, times, t_start, t_end, dt, derivative, cap=None):
    """E4. Returns (samples[n,D], sampling_times[n], segment_idx[n]) or None if out of range."""
    coeffs = _f64(coeffs)
    times = _f64(times)
    K, D, N = coeffs.shape
    if cap is None:
        # the reference's loop counter restarts at the start of the segment that
        # holds t_start (TRAJ_C:110-114), so size by t_end, not t_end - t_start
        cap = int(max(0.0, t_end) / dt) + 8
    out = np.zeros((cap, D))
    st = np.zeros(cap)
    seg = np.zeros(cap, dtype=np.int32)
    n = lib().mtgo_traj_evaluate_range(N, D, K, _d(coeffs), _d(times), C.c_double(t_start),
                                       C.c_double(t_end), C.c_double(dt), derivative, cap,
                                       _d(out), _d(st), seg.ctypes.data_as(_i32p))
    if n < 0:
        return None
    return out[:n].copy(), st[:n].copy(), seg[:n].copy()


# ------------------------------------------------------------------ extrema
def find_roots_jenkins_traub(coeffs_increasing):
    c = _f64(coeffs_increasing)
    re = np.zeros(128)
    im = np.zeros(128)
    n = C.c_int(0)
    ok = lib().mtgo_find_roots_jenkins_traub(_d(c), c.size, _d(re), _d(im), C.byref(n))
    if ok == -2:
        raise RuntimeError("oracle built without the reference rpoly (oracle/_ref missing)")
    return bool(ok), re[: n.value] + 1j * im[: n.value]


def poly_compute_min_max(c, t_start, t_end, derivative):
    c = _f64(c)
    v = [C.c_double(0.0) for _ in range(4)]
    rc = lib().mtgo_poly_compute_min_max(c.size, _d(c), C.c_double(t_start), C.c_double(t_end),
                                         derivative, *[C.byref(x) for x in v])
    if rc:
        return None
    return (v[0].value, v[1].value), (v[2].value, v[3].value)


def segment_candidate_times(seg_coeffs, derivative, t_start, t_end, dims=None):
    seg = _f64(seg_coeffs)
    D, N = seg.shape
    dims = np.arange(D, dtype=np.int32) if dims is None else np.ascontiguousarray(dims, dtype=np.int32)
    cand = np.zeros(128)
    n = lib().mtgo_segment_candidate_times(N, D, _d(seg), derivative, C.c_double(t_start),
                                           C.c_double(t_end), dims.ctypes.data_as(_ip), dims.size,
                                           _d(cand), 128)
    return None if n < 0 else cand[:n].copy()


def traj_min_max_magnitude(coeffs, times, derivative, dims=None):
    coeffs = _f64(coeffs)
    times = _f64(times)
    K, D, N = coeffs.shape
    dims = np.arange(D, dtype=np.int32) if dims is None else np.ascontiguousarray(dims, dtype=np.int32)
    mt, mv, Mt, Mv = (C.c_double(0.0) for _ in range(4))
    ms, Ms = C.c_int(0), C.c_int(0)
    rc = lib().mtgo_traj_min_max_magnitude(N, D, K, _d(coeffs), _d(times), derivative,
                                           dims.ctypes.data_as(_ip), dims.size, C.byref(mt),
                                           C.byref(mv), C.byref(ms), C.byref(Mt), C.byref(Mv),
                                           C.byref(Ms))
    if rc:
        return None
    return (mt.value, mv.value, ms.value), (Mt.value, Mv.value, Ms.value)


def opt_max_magnitude(coeffs, times, derivative):
    coeffs = _f64(coeffs)
    times = _f64(times)
    K, D, N = coeffs.shape
    t, v, s = C.c_double(0.0), C.c_double(0.0), C.c_int(0)
    rc = lib().mtgo_opt_max_magnitude(N, D, K, _d(coeffs), _d(times), derivative, C.byref(t),
                                      C.byref(v), C.byref(s))
    if rc:
        return None
    return t.value, v.value, s.value


def sampled_maximum_magnitude(coeffs, times, derivative, dt=0.01) -> float:
    coeffs = _f64(coeffs)
    times = _f64(times)
    K, D, N = coeffs.shape
    return lib().mtgo_sampled_maximum_magnitude(N, D, K, _d(coeffs), _d(times), derivative,
                                                C.c_double(dt))


def cost_numeric(coeffs, times, derivative, dt=0.001) -> float:
    coeffs = _f64(coeffs)
    times = _f64(times)
    K, D, N = coeffs.shape
    return lib().mtgo_cost_numeric(N, D, K, _d(coeffs), _d(times), derivative, C.c_double(dt))


# --------------------------------------------------------------------- tube
def tube_geometry(positions, radii) -> np.ndarray:
    positions = _f64(positions)
    radii = _f64(radii)
    K = positions.shape[0] - 1
    geom = np.zeros((K, 24))
    lib().mtgo_tube_geometry(K, _d(positions), _d(radii), _d(geom))
    return geom


def tube_flags(geom_seg, vertex_end, x) -> int:
    g = _f64(geom_seg)
    return lib().mtgo_tube_flags(_d(g), _d(_f64(vertex_end)), _d(_f64(x)))


def feasibility_sweep(coeffs, times, positions, radii, v_max, a_max, t_start, t_end, dt, cap=None):
    coeffs = _f64(coeffs)
    times = _f64(times)
    positions = _f64(positions)
    K, D, N = coeffs.shape
    assert D == 3
    if cap is None:
        cap = int(max(0.0, t_end) / dt) + 8
    pos = np.zeros((cap, 3))
    flags = np.zeros(cap, dtype=np.uint8)
    mv, ma = C.c_double(0.0), C.c_double(0.0)
    rad = _f64(radii) if radii is not None else None
    n = lib().mtgo_feasibility_sweep(N, K, _d(coeffs), _d(times), _d(positions),
                                     _d(rad) if rad is not None else None, C.c_double(v_max),
                                     C.c_double(a_max), C.c_double(t_start), C.c_double(t_end),
                                     C.c_double(dt), cap, _d(pos), flags.ctypes.data_as(_u8p),
                                     C.byref(mv), C.byref(ma))
    if n < 0:
        return None
    return pos[:n].copy(), flags[:n].copy(), mv.value, ma.value


# ------------------------------------------------------------------ N2 control points
def inverse_control_point_mapping(N: int, T: float) -> np.ndarray:
    out = np.zeros((N, N))
    if lib().mtgo_inverse_control_point_mapping(N, C.c_double(T), _d(out)):
        raise RuntimeError("singular control-point mapping")
    return out


def control_point_constraints(derivatives, times, positions=None, radii=None, N=10):
    """derivatives [K+1,h,D] -> dict(control_points [K,N,D], tube/cap_start/cap_end [K,N-2], sphere [K])."""
    derivatives = _f64(derivatives)
    times = _f64(times)
    K = times.shape[0]
    D = derivatives.shape[-1]
    cps = np.zeros((K, N, D))
    con = positions is not None and radii is not None and D == 3
    tube, cs, ce, sph = np.zeros((K, N - 2)), np.zeros((K, N - 2)), np.zeros((K, N - 2)), np.zeros(K)
    pos = _f64(positions) if con else None
    rad = _f64(radii) if con else None
    rc = lib().mtgo_control_point_constraints(N, K, D, _d(derivatives), _d(times), _d(pos) if con else None,
                                              _d(rad) if con else None, _d(cps), _d(tube) if con else None,
                                              _d(cs) if con else None, _d(ce) if con else None,
                                              _d(sph) if con else None)
    if rc:
        raise RuntimeError("control_point_constraints failed")
    out = dict(control_points=cps)
    if con:
        out.update(tube=tube, cap_start=cs, cap_end=ce, sphere=sph)
    return out


# ------------------------------------------------------------------ N3 composition / dump
def vertex_at_time(coeffs, times, t, max_derivative_order):
    """Trajectory::getVertexAtTime (TRAJ_C:248-254): evaluate(t, k) for k = 0..max -> [max+1, D]."""
    return np.stack([traj_evaluate(coeffs, times, t, k)[0] for k in range(max_derivative_order + 1)])


def sample_dump(coeffs, times, dt, max_rows):
    coeffs = _f64(coeffs)
    times = _f64(times)
    K, D, N = coeffs.shape
    rows = np.zeros((max_rows, 5 * D + 2))
    n = lib().mtgo_sample_dump(N, D, K, _d(coeffs), _d(times), C.c_double(dt), max_rows, _d(rows))
    return rows, n


# ------------------------------------------------------------------ N1 non-linear objective
def cost_gradient_derivative(N, derivative, times, mask, values, d_p):
    """NL_I:1537-1606 -> (J_d, grad [D, n_free])."""
    times = _f64(times)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    values = _f64(values)
    d_p = _f64(d_p)
    D = values.shape[2]
    J = C.c_double(0.0)
    g = np.zeros(d_p.size)
    rc = lib().mtgo_cost_gradient_derivative(N, D, times.size, derivative, _d(times), mask.ctypes.data_as(_u8p),
                                             _d(values), _d(d_p), C.byref(J), _d(g))
    if rc:
        raise ValueError(rc)
    return J.value, g.reshape(D, -1)


def soft_constraint_gradient(N, derivative, times, mask, values, d_p, ders, limits, weight, max_cost, increment,
                             central=True, want_grad=True):
    """NL_I:2735-2766, 2365-2490 -> (J_sc, grad [D, n_free] or None)."""
    times = _f64(times)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    values = _f64(values)
    d_p = _f64(d_p)
    D = values.shape[2]
    J = C.c_double(0.0)
    g = np.zeros(d_p.size)
    da = np.ascontiguousarray(ders, dtype=np.int32)
    la = _f64(limits)
    rc = lib().mtgo_soft_constraint_gradient(N, D, times.size, derivative, _d(times), mask.ctypes.data_as(_u8p),
                                             _d(values), _d(d_p), len(da), da.ctypes.data_as(_ip), _d(la),
                                             C.c_double(weight), C.c_double(max_cost), C.c_double(increment),
                                             1 if central else 0, C.byref(J), _d(g) if want_grad else None)
    if rc:
        raise ValueError(rc)
    return J.value, (g.reshape(D, -1) if want_grad else None)


# ------------------------------------------------------------------ N4 collision potential
def collision_cost(coeffs, times, grid, origin, res, min_bound, max_bound, dt, epsilon=0.5, robot_radius=0.5,
                   multiplier=1.0, want_grad=True):
    """NL_I:1608-1780 on a dense distance grid -> (J_c, grad [3, (K-1)*NF] or None, in_collision, n_checks)."""
    coeffs = _f64(coeffs)
    times = _f64(times)
    grid = _f64(grid)
    K, D, N = coeffs.shape
    assert D == 3
    size = np.ascontiguousarray(grid.shape, dtype=np.int32)
    org = np.ascontiguousarray(origin, dtype=np.int32)
    J = C.c_double(0.0)
    g = np.zeros(3 * (K - 1) * (N // 2 - 1))
    col, chk = C.c_int(0), C.c_int(0)
    lib().mtgo_collision_cost(N, K, _d(coeffs), _d(times), _d(grid), size.ctypes.data_as(_ip), org.ctypes.data_as(_ip),
                              C.c_double(res), _d(_f64(min_bound)), _d(_f64(max_bound)), C.c_double(dt),
                              C.c_double(epsilon), C.c_double(robot_radius), C.c_double(multiplier), C.byref(J),
                              _d(g) if want_grad else None, C.byref(col), C.byref(chk))
    return J.value, (g.reshape(3, -1) if want_grad else None), bool(col.value), chk.value


# ------------------------------------------------------------------ G1 device generator: host twin (numpy)
def _philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on uint64-held 32-bit lanes (vectorised over candidates)."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return c0, c1, c2, c3


def generate_candidates(B, K, D, seed, first_index=0, pos_min=-10.0, pos_max=10.0, v_max=3.0, a_max=5.0, magic=6.5):
    """Host twin of mtg_generate_candidates_batch: createRandomVertices' rejection rule (VTX_C:65-72) and Nfabian
    times (VTX_C:252-269) on Philox4x32-10 keyed by (seed; candidate index, draw, block).
    Returns positions [B, K+1, D], times [B, K]."""
    lo = np.broadcast_to(np.asarray(pos_min, dtype=np.float64), (D,))
    hi = np.broadcast_to(np.asarray(pos_max, dtype=np.float64), (D,))
    idx = (np.arange(B, dtype=np.uint64) + np.uint64(first_index))
    g0, g1 = idx & np.uint64(0xFFFFFFFF), idx >> np.uint64(32)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    draw = np.zeros(B, dtype=np.uint64)
    pos = np.zeros((B, K + 1, D))
    times = np.zeros((B, K))
    last = np.zeros((B, D))

    def u53(a, b):
        return (((a >> np.uint64(5)) << np.uint64(26)) | (b >> np.uint64(6))).astype(np.float64) * (1.0 / 9007199254740992.0)

    for v in range(K + 1):
        todo = np.ones(B, dtype=bool)
        cur = np.zeros((B, D))
        dist = np.zeros(B)
        while todo.any():
            sel = np.flatnonzero(todo)
            z = np.zeros(sel.size, dtype=np.uint64)
            c = _philox4x32_10(g0[sel], g1[sel], draw[sel], z, k0, k1)
            u = [u53(c[0], c[1]), u53(c[2], c[3])]
            if D > 2:
                e = _philox4x32_10(g0[sel], g1[sel], draw[sel], z + np.uint64(1), k0, k1)
                u += [u53(e[0], e[1]), u53(e[2], e[3])]
            draw[sel] += np.uint64(1)
            s = np.zeros(sel.size)
            for dim in range(D):
                x = lo[dim] + u[dim] * (hi[dim] - lo[dim])
                cur[sel, dim] = x
                dd = x - last[sel, dim]
                s = s + dd * dd
            d_ = np.sqrt(s)
            dist[sel] = d_
            ok = np.ones(sel.size, dtype=bool) if v == 0 else d_ > 0.2
            todo[sel[ok]] = False
        pos[:, v] = cur
        last = cur.copy()
        if v >= 1:
            times[:, v - 1] = dist / v_max * 2 * (1.0 + magic * v_max / a_max * np.exp(-dist / v_max * 2))
    return pos, times

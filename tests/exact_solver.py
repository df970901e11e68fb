"""TEST INFRASTRUCTURE — 60-digit mpmath solve of the SAME normal equations the
reference builds (A, Q, C, R = C^T A^-T Q A^-1 C, d_p = -R_pp^-1 R_pf d_f;
LIN_I:101-111, 171-252, 306-379, 557-573). It is the arbiter between two fp64
evaluation orders: the oracle's dense-QR restatement and the
CUDA path are both compared against it norm-wise.
"""
import mpmath as mp
import numpy as np

mp.mp.dps = 60


def _base(n, i):
    # i!/(i-n)!  (POLY_C:145-161)
    if i < n:
        return mp.mpf(0)
    r = mp.mpf(1)
    for q in range(i - n + 1, i + 1):
        r *= q
    return r


def constraint_columns(mask):
    """Row->column map of the reordering matrix C (LIN_I:171-252)."""
    Kp1, h = mask.shape
    K = Kp1 - 1
    all_c, fixed, free = [], [], []
    for v in range(Kp1):
        occ = 1 if v in (0, K) else 2
        for co in range(occ):
            for k in range(h):
                all_c.append((v, k))
                if co == 0:
                    (fixed if mask[v, k] else free).append((v, k))
    cols = {c: i for i, c in enumerate(fixed)}
    cols.update({c: len(fixed) + i for i, c in enumerate(free)})
    return all_c, fixed, free, cols


def exact_solve(N, derivative, times, mask, values):
    """Returns (coeffs[K,D,N] float64-rounded, cost, d_p[D,n_free])."""
    d = derivative
    h = N // 2
    K = len(times)
    D = values.shape[2]
    all_c, fixed, free, cols = constraint_columns(mask)
    nf, npp = len(fixed), len(free)
    n = nf + npp
    R = mp.zeros(n, n)
    Ainvs = []
    for i in range(K):
        t = mp.mpf(float(times[i]))
        A = mp.zeros(N, N)
        for r in range(h):
            A[r, r] = _base(r, r)
            for j in range(r, N):
                A[r + h, j] = _base(r, j) * t ** (j - r)
        Q = mp.zeros(N, N)
        for a in range(d, N):
            for b in range(d, N):
                e = a + b - 2 * d + 1
                Q[a, b] = _base(d, a) * _base(d, b) * t ** e * 2 / e
        Ai = A ** -1
        Ainvs.append(Ai)
        H = Ai.T * Q * Ai
        for r in range(N):
            for c in range(N):
                R[cols[all_c[i * N + r]], cols[all_c[i * N + c]]] += H[r, c]
    coeffs = np.zeros((K, D, N))
    d_p = np.zeros((D, npp))
    cost = mp.mpf(0)
    for dim in range(D):
        df = mp.matrix([mp.mpf(float(values[v, k, dim])) for (v, k) in fixed])
        if npp:
            dp = mp.lu_solve(R[nf:, nf:], -R[nf:, :nf] * df)
            d_p[dim] = [float(x) for x in dp]
            dall = mp.matrix(list(df) + list(dp))
        else:
            dall = df
        cost += (dall.T * R * dall)[0] / 2
        for i in range(K):
            nd = mp.matrix([dall[cols[all_c[i * N + r]]] for r in range(N)])
            c = Ainvs[i] * nd
            coeffs[i, dim, :] = [float(x) for x in c]
    return coeffs, float(cost), d_p


def exact_cost_derivative(N, derivative, times, mask, values, d_p):
    """J_d = sum_dim [d_f; d_p]^T R(times) [d_f; d_p] (NL_I:1537-1606, no 1/2) in 60 digits,
    for the float64 inputs as given (d_p [D, n_free] held fixed)."""
    d = derivative
    h = N // 2
    K = len(times)
    D = values.shape[2]
    all_c, fixed, free, cols = constraint_columns(mask)
    nf = len(fixed)
    J = mp.mpf(0)
    for i in range(K):
        t = mp.mpf(float(times[i]))
        A = mp.zeros(N, N)
        for r in range(h):
            A[r, r] = _base(r, r)
            for j in range(r, N):
                A[r + h, j] = _base(r, j) * t ** (j - r)
        Q = mp.zeros(N, N)
        for a in range(d, N):
            for b in range(d, N):
                e = a + b - 2 * d + 1
                Q[a, b] = _base(d, a) * _base(d, b) * t ** e * 2 / e
        for dim in range(D):
            dv = []
            for r in range(N):
                v, k = all_c[i * N + r]
                col = cols[(v, k)]
                dv.append(mp.mpf(float(values[v, k, dim])) if col < nf else mp.mpf(float(d_p[dim][col - nf])))
            c = mp.lu_solve(A, mp.matrix(dv))
            J += (c.T * Q * c)[0]
    return J

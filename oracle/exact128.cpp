// TEST INFRASTRUCTURE — binary128 arbiter of the linear solve. Never shipped, never on the product path.
//
// The reference's own fp64 evaluation order (A(T) inverted numerically, A^-T Q A^-1, SparseQR) is itself
// 1e-11 .. 3e-7 away from the exact solution of its normal equations, depending on the cost derivative and
// the segment times (tests/test_oracle.py::test_oracle_vs_mpmath). To hold EVERY item of a batch to the
// north-star's 1e-9 — not only the few that the slow 60-digit mpmath solve (tests/exact_solver.py) can
// afford — the same normal equations are solved here in IEEE binary128 (113-bit significand, ~34 digits):
// A and Q as written in LIN_I:101-111 / 557-573, a pivoted inverse of A, H = A^-T Q A^-1, the reordering of
// LIN_I:171-252, R = C^T H C, d_p = -R_pp^-1 R_pf d_f (LIN_I:337-379), coefficients (LIN_I:254-275) and
// cost (LIN_I:113-130). With cond(A) <= 1e12 the rounded result is correct to the last double digit;
// tests/test_oracle.py pins it against the mpmath solve.
#include <cstddef>
#include <cstdint>
#include <vector>

namespace {
typedef __float128 q128;
typedef std::vector<q128> QVec;

q128 q_abs(q128 x) { return x < 0 ? -x : x; }

q128 q_sqrt(q128 s) {
  if (!(s > 0)) return 0;
  q128 r = static_cast<q128>(__builtin_sqrt(static_cast<double>(s)));
  for (int it = 0; it < 4; ++it) r = (r + s / r) / 2;  // Newton from the double root: 53 -> 113 bits in 2 steps
  return r;
}

q128 base_coeff(int n, int i) {  // i!/(i-n)!, POLY_C:145-161
  if (i < n) return 0;
  q128 r = 1;
  for (int q = i - n + 1; q <= i; ++q) r *= q;
  return r;
}

bool invert(int n, const QVec& M, QVec* inv) {
  QVec a(M);
  inv->assign(static_cast<size_t>(n) * n, 0);
  for (int i = 0; i < n; ++i) (*inv)[i * n + i] = 1;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (q_abs(a[r * n + c]) > q_abs(a[piv * n + c])) piv = r;
    if (a[piv * n + c] == 0) return false;
    if (piv != c)
      for (int j = 0; j < n; ++j) {
        std::swap(a[c * n + j], a[piv * n + j]);
        std::swap((*inv)[c * n + j], (*inv)[piv * n + j]);
      }
    const q128 d = 1 / a[c * n + c];
    for (int j = 0; j < n; ++j) {
      a[c * n + j] *= d;
      (*inv)[c * n + j] *= d;
    }
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const q128 f = a[r * n + c];
      if (f == 0) continue;
      for (int j = 0; j < n; ++j) {
        a[r * n + j] -= f * a[c * n + j];
        (*inv)[r * n + j] -= f * (*inv)[c * n + j];
      }
    }
  }
  return true;
}

int solve_one(int N, int D, int K, int derivative, const double* times, const uint8_t* mask, const double* values,
              double* coeffs, double* cost, double* d_p_out) {
  const int h = N / 2, d = derivative;
  // column of every (vertex, derivative): fixed first, then free, each ordered by (v, k)  (LIN_H:289-296)
  std::vector<int> col((K + 1) * h, -1);
  int nf = 0, np = 0;
  for (int i = 0; i < (K + 1) * h; ++i)
    if (mask[i]) col[i] = nf++;
  for (int i = 0; i < (K + 1) * h; ++i)
    if (!mask[i]) col[i] = nf + np++;
  const int n = nf + np;
  QVec R(static_cast<size_t>(n) * n, 0), Ainv_all(static_cast<size_t>(K) * N * N), Q_all(static_cast<size_t>(K) * N * N, 0);
  for (int i = 0; i < K; ++i) {
    const q128 t = times[i];
    if (!(times[i] > 0.0)) return -3;
    QVec tp(2 * N + 2);
    tp[0] = 1;
    for (size_t j = 1; j < tp.size(); ++j) tp[j] = tp[j - 1] * t;
    QVec A(static_cast<size_t>(N) * N, 0), Ai;
    for (int r = 0; r < h; ++r) {
      A[r * N + r] = base_coeff(r, r);
      for (int j = r; j < N; ++j) A[(r + h) * N + j] = base_coeff(r, j) * tp[j - r];
    }
    if (!invert(N, A, &Ai)) return -4;
    q128* Q = &Q_all[static_cast<size_t>(i) * N * N];
    for (int a = d; a < N; ++a)
      for (int b = d; b < N; ++b) {
        const int e = a + b - 2 * d + 1;
        Q[a * N + b] = base_coeff(d, a) * base_coeff(d, b) * tp[e] * 2 / e;
      }
    for (int j = 0; j < N * N; ++j) Ainv_all[static_cast<size_t>(i) * N * N + j] = Ai[j];
    // H = Ai^T Q Ai, scattered into R through the reordering
    QVec QA(static_cast<size_t>(N) * N, 0), H(static_cast<size_t>(N) * N, 0);
    for (int r = 0; r < N; ++r)
      for (int c = 0; c < N; ++c) {
        q128 s = 0;
        for (int k = 0; k < N; ++k) s += Q[r * N + k] * Ai[k * N + c];
        QA[r * N + c] = s;
      }
    for (int r = 0; r < N; ++r)
      for (int c = 0; c < N; ++c) {
        q128 s = 0;
        for (int k = 0; k < N; ++k) s += Ai[k * N + r] * QA[k * N + c];
        H[r * N + c] = s;
      }
    auto column = [&](int local) { return col[(i + (local >= h ? 1 : 0)) * h + local % h]; };
    for (int r = 0; r < N; ++r)
      for (int c = 0; c < N; ++c) R[static_cast<size_t>(column(r)) * n + column(c)] += H[r * N + c];
  }
  // d_all per dimension; solve R_pp d_p = -R_pf d_f by Cholesky of the (symmetric positive definite) R_pp
  QVec L(static_cast<size_t>(np) * np, 0);
  for (int j = 0; j < np; ++j) {
    q128 s = R[static_cast<size_t>(nf + j) * n + nf + j];
    for (int k = 0; k < j; ++k) s -= L[j * np + k] * L[j * np + k];
    if (!(s > 0)) return -4;
    const q128 ljj = q_sqrt(s);
    L[j * np + j] = ljj;
    for (int i = j + 1; i < np; ++i) {
      q128 v = (R[static_cast<size_t>(nf + i) * n + nf + j] + R[static_cast<size_t>(nf + j) * n + nf + i]) / 2;
      for (int k = 0; k < j; ++k) v -= L[i * np + k] * L[j * np + k];
      L[i * np + j] = v / ljj;
    }
  }
  q128 total = 0;
  for (int dim = 0; dim < D; ++dim) {
    QVec dall(n, 0);
    for (int i = 0; i < (K + 1) * h; ++i)
      if (mask[i]) dall[col[i]] = values[static_cast<size_t>(i) * D + dim];
    QVec y(np);
    for (int r = 0; r < np; ++r) {
      q128 s = 0;
      for (int c = 0; c < nf; ++c) s -= R[static_cast<size_t>(nf + r) * n + c] * dall[c];
      y[r] = s;
    }
    for (int i = 0; i < np; ++i) {
      q128 s = y[i];
      for (int k = 0; k < i; ++k) s -= L[i * np + k] * y[k];
      y[i] = s / L[i * np + i];
    }
    for (int i = np - 1; i >= 0; --i) {
      q128 s = y[i];
      for (int k = i + 1; k < np; ++k) s -= L[k * np + i] * y[k];
      y[i] = s / L[i * np + i];
    }
    for (int r = 0; r < np; ++r) {
      dall[nf + r] = y[r];
      if (d_p_out) d_p_out[static_cast<size_t>(dim) * np + r] = static_cast<double>(y[r]);
    }
    for (int i = 0; i < K; ++i) {
      const q128* Ai = &Ainv_all[static_cast<size_t>(i) * N * N];
      const q128* Q = &Q_all[static_cast<size_t>(i) * N * N];
      QVec c(N);
      for (int r = 0; r < N; ++r) {
        q128 s = 0;
        for (int k = 0; k < N; ++k) s += Ai[r * N + k] * dall[col[(i + (k >= h ? 1 : 0)) * h + k % h]];
        c[r] = s;
        coeffs[(static_cast<size_t>(i) * D + dim) * N + r] = static_cast<double>(s);
      }
      for (int r = d; r < N; ++r)
        for (int k = d; k < N; ++k) total += c[r] * Q[r * N + k] * c[k];
    }
  }
  if (cost) *cost = static_cast<double>(total / 2);
  return 0;
}
}  // namespace

extern "C" {
// One problem: mask [(K+1)*h], values [(K+1)*h*D] -> coeffs [K*D*N], cost, d_p [D*n_free] (or NULL)
int mtgo_solve_exact128(int N, int D, int K, int derivative, const double* times, const uint8_t* mask,
                        const double* values, double* coeffs, double* cost, double* d_p) {
  return solve_one(N, D, K, derivative, times, mask, values, coeffs, cost, d_p);
}

// B problems sharing one mask (OpenMP over the batch): times [B][K], values [B][(K+1)*h*D]
int mtgo_solve_exact128_batch(int B, int N, int D, int K, int derivative, const double* times, const uint8_t* mask,
                              const double* values, double* coeffs, double* cost, double* d_p, int n_free,
                              int n_threads) {
  const int h = N / 2;
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads > 0 ? n_threads : 1) reduction(+ : bad)
  for (int b = 0; b < B; ++b) {
    const int rc = solve_one(N, D, K, derivative, times + static_cast<size_t>(b) * K, mask,
                             values + static_cast<size_t>(b) * (K + 1) * h * D,
                             coeffs + static_cast<size_t>(b) * K * D * N, cost ? cost + b : nullptr,
                             d_p ? d_p + static_cast<size_t>(b) * D * n_free : nullptr);
    if (rc) ++bad;
  }
  return bad;
}
}

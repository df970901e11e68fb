// TEST INFRASTRUCTURE — CPU oracle (see oracle.h for scope, pinning and the
// citation tags). Plain C++14, single-threaded unless stated, compiled with
// -ffp-contract=off so that its arithmetic is the written IEEE sequence.
#include "oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <random>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef MTG_ORACLE_NO_RPOLY
extern "C" void mtg_ref_rpoly(double* coefficients_decreasing, int* degree,
                              double* roots_real, double* roots_imag);
#endif

namespace {

// ---------------------------------------------------------------- P2 table
// POLY_C:145-161: row 0 all ones; row n = (i-(n-1)) * row n-1  ==> i!/(i-n)!.
struct BaseTable {
  double v[MTGO_MAX_CONV][MTGO_MAX_CONV];
  BaseTable() {
    const int n_tab = MTGO_MAX_CONV;
    for (int n = 0; n < n_tab; ++n)
      for (int i = 0; i < n_tab; ++i) v[n][i] = 0.0;
    for (int i = 0; i < n_tab; ++i) v[0][i] = 1.0;
    const int deg = n_tab - 1;
    int order = deg;
    for (int n = 1; n < n_tab; ++n) {
      for (int i = deg - order; i < n_tab; ++i)
        v[n][i] = static_cast<double>(order - deg + i) * v[n - 1][i];
      --order;
    }
  }
};
const BaseTable kBase;
inline double Bc(int derivative, int j) { return kBase.v[derivative][j]; }

// ------------------------------------------------------- small dense algebra
typedef std::vector<double> Vec;

// C = A(m x k) * B(k x n), row-major, inner index ascending.
void matmul(const double* A, const double* B, double* C, int m, int k, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int p = 0; p < k; ++p) s += A[i * k + p] * B[p * n + j];
      C[i * n + j] = s;
    }
}

// Partial-pivot Gauss-Jordan inverse (what Eigen's fixed-size inverse() does
// for sizes > 4, up to rounding order). Returns false if singular.
bool general_inverse(int n, const double* M, double* Minv) {
  Vec a(M, M + n * n);
  Vec b(static_cast<size_t>(n) * n, 0.0);
  for (int i = 0; i < n; ++i) b[i * n + i] = 1.0;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = std::fabs(a[c * n + c]);
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(a[r * n + c]) > best) {
        best = std::fabs(a[r * n + c]);
        piv = r;
      }
    if (best == 0.0) return false;
    if (piv != c)
      for (int j = 0; j < n; ++j) {
        std::swap(a[c * n + j], a[piv * n + j]);
        std::swap(b[c * n + j], b[piv * n + j]);
      }
    const double inv = 1.0 / a[c * n + c];
    for (int j = 0; j < n; ++j) {
      a[c * n + j] *= inv;
      b[c * n + j] *= inv;
    }
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = a[r * n + c];
      if (f == 0.0) continue;
      for (int j = 0; j < n; ++j) {
        a[r * n + j] -= f * a[c * n + j];
        b[r * n + j] -= f * b[c * n + j];
      }
    }
  }
  std::memcpy(Minv, b.data(), sizeof(double) * n * n);
  return true;
}

// Dense Householder QR solve of the square system M x = rhs (n_rhs columns,
// rhs row-major [n][n_rhs]); the closest dense stand-in for the reference's
// Eigen::SparseQR factor/solve (LIN_I:364-374).
bool qr_solve(int n, const double* M, int n_rhs, double* rhs) {
  Vec a(M, M + n * n);
  Vec v(n);
  for (int c = 0; c < n; ++c) {
    double norm = 0.0;
    for (int r = c; r < n; ++r) norm += a[r * n + c] * a[r * n + c];
    norm = std::sqrt(norm);
    if (norm == 0.0) return false;
    const double alpha = (a[c * n + c] > 0.0) ? -norm : norm;
    for (int r = c; r < n; ++r) v[r] = a[r * n + c];
    v[c] -= alpha;
    double vtv = 0.0;
    for (int r = c; r < n; ++r) vtv += v[r] * v[r];
    if (vtv == 0.0) continue;
    const double beta = 2.0 / vtv;
    for (int j = c; j < n; ++j) {
      double s = 0.0;
      for (int r = c; r < n; ++r) s += v[r] * a[r * n + j];
      s *= beta;
      for (int r = c; r < n; ++r) a[r * n + j] -= s * v[r];
    }
    for (int j = 0; j < n_rhs; ++j) {
      double s = 0.0;
      for (int r = c; r < n; ++r) s += v[r] * rhs[r * n_rhs + j];
      s *= beta;
      for (int r = c; r < n; ++r) rhs[r * n_rhs + j] -= s * v[r];
    }
  }
  for (int j = 0; j < n_rhs; ++j)
    for (int r = n - 1; r >= 0; --r) {
      double s = rhs[r * n_rhs + j];
      for (int c = r + 1; c < n; ++c) s -= a[r * n + c] * rhs[c * n_rhs + j];
      if (a[r * n + r] == 0.0) return false;
      rhs[r * n_rhs + j] = s / a[r * n + r];
    }
  return true;
}

bool cholesky_solve(int n, const double* M, int n_rhs, double* rhs) {
  Vec l(static_cast<size_t>(n) * n, 0.0);
  for (int j = 0; j < n; ++j) {
    double d = M[j * n + j];
    for (int k = 0; k < j; ++k) d -= l[j * n + k] * l[j * n + k];
    if (!(d > 0.0)) return false;
    const double ljj = std::sqrt(d);
    l[j * n + j] = ljj;
    for (int i = j + 1; i < n; ++i) {
      // symmetrise: the reference's R_pp is symmetric only up to rounding
      double s = 0.5 * (M[i * n + j] + M[j * n + i]);
      for (int k = 0; k < j; ++k) s -= l[i * n + k] * l[j * n + k];
      l[i * n + j] = s / ljj;
    }
  }
  for (int c = 0; c < n_rhs; ++c) {
    for (int i = 0; i < n; ++i) {
      double s = rhs[i * n_rhs + c];
      for (int k = 0; k < i; ++k) s -= l[i * n + k] * rhs[k * n_rhs + c];
      rhs[i * n_rhs + c] = s / l[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = rhs[i * n_rhs + c];
      for (int k = i + 1; k < n; ++k) s -= l[k * n + i] * rhs[k * n_rhs + c];
      rhs[i * n_rhs + c] = s / l[i * n + i];
    }
  }
  return true;
}

// ------------------------------------------------------------- P3, P4, P5
void quadratic_cost_jacobian(int N, int derivative, double t, double* Q) {
  for (int i = 0; i < N * N; ++i) Q[i] = 0.0;
  for (int col = 0; col < N - derivative; ++col)
    for (int row = 0; row < N - derivative; ++row) {
      const double exponent = (N - 1 - derivative) * 2 + 1 - row - col;
      Q[(N - 1 - row) * N + (N - 1 - col)] =
          Bc(derivative, N - 1 - row) * Bc(derivative, N - 1 - col) *
          std::pow(t, exponent) * 2.0 / exponent;
    }
}

void base_coeffs_with_time(int N, int derivative, double t, double* out) {
  for (int j = 0; j < N; ++j) out[j] = 0.0;
  out[derivative] = Bc(derivative, derivative);
  if (std::fabs(t) < std::numeric_limits<double>::epsilon()) return;
  double t_power = t;
  for (int j = derivative + 1; j < N; ++j) {
    out[j] = Bc(derivative, j) * t_power;
    t_power = t_power * t;
  }
}

void setup_mapping_matrix(int N, double t, double* A) {
  const int h = N / 2;
  for (int i = 0; i < h; ++i) {
    base_coeffs_with_time(N, i, 0.0, A + i * N);
    base_coeffs_with_time(N, i, t, A + (i + h) * N);
  }
}

void invert_mapping_matrix(int N, const double* A, double* Ainv) {
  const int h = N / 2;
  Vec a_inv(static_cast<size_t>(h) * h, 0.0), c(h * h), d(h * h), d_inv(h * h);
  for (int i = 0; i < h; ++i) a_inv[i * h + i] = 1.0 / A[i * N + i];
  for (int i = 0; i < h; ++i)
    for (int j = 0; j < h; ++j) {
      c[i * h + j] = A[(i + h) * N + j];
      d[i * h + j] = A[(i + h) * N + (j + h)];
    }
  general_inverse(h, d.data(), d_inv.data());
  Vec neg(h * h), t1(h * h), t2(h * h);
  for (int i = 0; i < h * h; ++i) neg[i] = -d_inv[i];
  matmul(neg.data(), c.data(), t1.data(), h, h, h);      // (-D^-1) C
  matmul(t1.data(), a_inv.data(), t2.data(), h, h, h);   // ... diag^-1
  for (int i = 0; i < h; ++i)
    for (int j = 0; j < h; ++j) {
      Ainv[i * N + j] = a_inv[i * h + j];
      Ainv[i * N + (j + h)] = 0.0;
      Ainv[(i + h) * N + j] = t2[i * h + j];
      Ainv[(i + h) * N + (j + h)] = d_inv[i * h + j];
    }
}

// -------------------------------------------------- the optimisation object
struct Problem {
  int N, D, K, derivative, h;
  const double* times;
  const uint8_t* mask;    // [(K+1)][h]
  const double* values;   // [(K+1)][h][D]
  // state
  int n_all, n_fixed, n_free;
  std::vector<int> col_of_row;      // reordering matrix C, one 1 per row
  Vec d_f;                          // [D][n_fixed]
  Vec d_p;                          // [D][n_free]
  Vec Q, Ainv;                      // [K][N][N]
};

bool has_constraint(const Problem& p, int v, int k) {
  return p.mask[v * p.h + k] != 0;
}

// P6 LIN_I:277-304 (debug prints of :287-292 excluded)
int update_segment_times(Problem& p, const double* times) {
  const int N = p.N;
  p.Q.resize(static_cast<size_t>(p.K) * N * N);
  p.Ainv.resize(static_cast<size_t>(p.K) * N * N);
  Vec A(static_cast<size_t>(N) * N);
  for (int i = 0; i < p.K; ++i) {
    const double T = times[i];
    if (!(T > 0.0)) return -3;  // CHECK_GT(segment_time, 0)
    quadratic_cost_jacobian(N, p.derivative, T, &p.Q[static_cast<size_t>(i) * N * N]);
    setup_mapping_matrix(N, T, A.data());
    invert_mapping_matrix(N, A.data(), &p.Ainv[static_cast<size_t>(i) * N * N]);
  }
  return 0;
}

// P7 LIN_I:171-252
void setup_constraint_reordering(Problem& p) {
  struct Key { int v, k; };
  std::vector<Key> all, fixed, free_;
  const int n_vertices = p.K + 1;
  for (int v = 0; v < n_vertices; ++v) {
    const int occ = (v == 0 || v == p.K) ? 1 : 2;
    for (int co = 0; co < occ; ++co)
      for (int k = 0; k < p.h; ++k) {
        all.push_back(Key{v, k});
        // std::set insert: unique, ordered by (vertex_idx, constraint_idx).
        // Enumeration is already in that order, so "insert if new" == append
        // on first occurrence.
        if (co == 0) (has_constraint(p, v, k) ? fixed : free_).push_back(Key{v, k});
      }
  }
  p.n_all = static_cast<int>(all.size());
  p.n_fixed = static_cast<int>(fixed.size());
  p.n_free = static_cast<int>(free_.size());
  p.col_of_row.assign(p.n_all, -1);
  p.d_f.assign(static_cast<size_t>(p.D) * p.n_fixed, 0.0);
  int row = 0;
  for (const Key& ca : all) {
    int col = 0;
    for (const Key& cf : fixed) {
      if (ca.v == cf.v && ca.k == cf.k) {
        p.col_of_row[row] = col;
        for (int d = 0; d < p.D; ++d)
          p.d_f[static_cast<size_t>(d) * p.n_fixed + col] =
              p.values[(static_cast<size_t>(cf.v) * p.h + cf.k) * p.D + d];
      }
      ++col;
    }
    for (const Key& cp : free_) {
      if (ca.v == cp.v && ca.k == cp.k) p.col_of_row[row] = col;
      ++col;
    }
    ++row;
  }
}

// P8a LIN_I:306-335 : H_i = Ainv^T Q Ainv ; R = C^T H C (scatter-add)
void construct_R(const Problem& p, Vec* R) {
  const int N = p.N, n = p.n_fixed + p.n_free;
  R->assign(static_cast<size_t>(n) * n, 0.0);
  Vec At(N * N), t1(N * N), H(N * N);
  for (int i = 0; i < p.K; ++i) {
    const double* Ai = &p.Ainv[static_cast<size_t>(i) * N * N];
    const double* Q = &p.Q[static_cast<size_t>(i) * N * N];
    for (int r = 0; r < N; ++r)
      for (int c = 0; c < N; ++c) At[r * N + c] = Ai[c * N + r];
    matmul(At.data(), Q, t1.data(), N, N, N);
    matmul(t1.data(), Ai, H.data(), N, N, N);
    for (int r = 0; r < N; ++r)
      for (int c = 0; c < N; ++c)
        (*R)[static_cast<size_t>(p.col_of_row[i * N + r]) * n + p.col_of_row[i * N + c]] +=
            H[r * N + c];
  }
}

// P8c LIN_I:254-275
void update_segments_from_compact(const Problem& p, double* coeffs) {
  const int N = p.N;
  Vec d_all(p.n_fixed + p.n_free), new_d(N);
  for (int dim = 0; dim < p.D; ++dim) {
    for (int c = 0; c < p.n_fixed; ++c) d_all[c] = p.d_f[static_cast<size_t>(dim) * p.n_fixed + c];
    for (int c = 0; c < p.n_free; ++c)
      d_all[p.n_fixed + c] = p.d_p[static_cast<size_t>(dim) * p.n_free + c];
    for (int i = 0; i < p.K; ++i) {
      for (int r = 0; r < N; ++r) new_d[r] = d_all[p.col_of_row[i * N + r]];
      const double* Ai = &p.Ainv[static_cast<size_t>(i) * N * N];
      double* out = coeffs + (static_cast<size_t>(i) * p.D + dim) * N;
      for (int r = 0; r < N; ++r) {
        double s = 0.0;
        for (int c = 0; c < N; ++c) s += Ai[r * N + c] * new_d[c];
        out[r] = s;
      }
    }
  }
}

// P8d LIN_I:113-130
double compute_cost(const Problem& p, const double* coeffs) {
  const int N = p.N;
  double cost = 0.0;
  Vec tmp(N);
  for (int i = 0; i < p.K; ++i) {
    const double* Q = &p.Q[static_cast<size_t>(i) * N * N];
    for (int dim = 0; dim < p.D; ++dim) {
      const double* c = coeffs + (static_cast<size_t>(i) * p.D + dim) * N;
      for (int col = 0; col < N; ++col) {  // (c^T Q)
        double s = 0.0;
        for (int r = 0; r < N; ++r) s += c[r] * Q[r * N + col];
        tmp[col] = s;
      }
      double partial = 0.0;
      for (int r = 0; r < N; ++r) partial += tmp[r] * c[r];
      cost += partial;
    }
  }
  return 0.5 * cost;
}

// P1 LIN_I:46-99 (validation + state) ; constraints above N/2-1 cannot be
// represented in mask[][h] and are dropped by the caller like LIN_I:74-95.
int setup_problem(Problem& p, int N, int D, int K, int derivative,
                  const double* times, const uint8_t* mask,
                  const double* values) {
  if (N < 2 || N > MTGO_MAX_N || (N % 2) != 0) return -1;
  if (derivative < 0 || derivative > N / 2 - 1) return -2;
  if (K < 1 || D < 1) return -1;
  p.N = N; p.D = D; p.K = K; p.derivative = derivative; p.h = N / 2;
  p.times = times; p.mask = mask; p.values = values;
  const int rc = update_segment_times(p, times);
  if (rc) return rc;
  setup_constraint_reordering(p);
  return 0;
}

// P8b LIN_I:337-379
int solve_linear(Problem& p, int solver, double* coeffs, Vec* R_keep) {
  p.d_p.assign(static_cast<size_t>(p.D) * p.n_free, 0.0);
  if (p.n_free == 0) {  // fully constrained shortcut LIN_I:342-348
    update_segments_from_compact(p, coeffs);
    if (R_keep) construct_R(p, R_keep);
    return 0;
  }
  Vec R;
  construct_R(p, &R);
  const int nf = p.n_fixed, np = p.n_free, n = nf + np;
  Vec Rpp(static_cast<size_t>(np) * np), rhs(static_cast<size_t>(np) * p.D);
  for (int r = 0; r < np; ++r)
    for (int c = 0; c < np; ++c) Rpp[static_cast<size_t>(r) * np + c] = R[static_cast<size_t>(nf + r) * n + nf + c];
  for (int dim = 0; dim < p.D; ++dim)
    for (int r = 0; r < np; ++r) {
      double s = 0.0;  // (-Rpf) * d_f
      for (int c = 0; c < nf; ++c)
        s += (-R[static_cast<size_t>(nf + r) * n + c]) * p.d_f[static_cast<size_t>(dim) * nf + c];
      rhs[static_cast<size_t>(r) * p.D + dim] = s;
    }
  const bool ok = (solver == 1) ? cholesky_solve(np, Rpp.data(), p.D, rhs.data())
                                : qr_solve(np, Rpp.data(), p.D, rhs.data());
  if (!ok) return -4;
  for (int dim = 0; dim < p.D; ++dim)
    for (int r = 0; r < np; ++r)
      p.d_p[static_cast<size_t>(dim) * np + r] = rhs[static_cast<size_t>(r) * p.D + dim];
  update_segments_from_compact(p, coeffs);
  if (R_keep) R_keep->swap(R);
  return 0;
}

// NL_I:1537-1606 : J_d summed over dimensions (no 1/2)
double cost_derivative(const Problem& p) {
  Vec R;
  construct_R(p, &R);
  const int nf = p.n_fixed, np = p.n_free, n = nf + np;
  double J = 0.0;
  for (int dim = 0; dim < p.D; ++dim) {
    const double* df = &p.d_f[static_cast<size_t>(dim) * nf];
    const double* dp = np ? &p.d_p[static_cast<size_t>(dim) * np] : nullptr;
    double t_ff = 0.0, t_fp = 0.0, t_pf = 0.0, t_pp = 0.0;
    for (int r = 0; r < nf; ++r) {
      double s = 0.0;
      for (int c = 0; c < nf; ++c) s += R[static_cast<size_t>(r) * n + c] * df[c];
      t_ff += df[r] * s;
    }
    for (int r = 0; r < np; ++r) {
      double s = 0.0, s2 = 0.0;
      for (int c = 0; c < nf; ++c) s += R[static_cast<size_t>(nf + r) * n + c] * df[c];
      for (int c = 0; c < np; ++c) s2 += R[static_cast<size_t>(nf + r) * n + nf + c] * dp[c];
      t_pf += dp[r] * s;
      t_pp += dp[r] * s2;
    }
    t_fp = t_pf;  // d_f^T R_pf^T d_p is the transpose of the same scalar
    J += t_ff + t_fp + t_pf + t_pp;
  }
  return J;
}

// ------------------------------------------------------------- E1..E4
double poly_evaluate(int N, const double* c, double t, int derivative) {
  if (derivative >= N) return 0.0;
  const int top = N - 1;
  double result = Bc(derivative, top) * c[top];
  for (int j = top - 1; j >= derivative; --j) {
    result *= t;
    result += Bc(derivative, j) * c[j];
  }
  return result;
}

void segment_evaluate(int N, int D, const double* seg_coeffs, double t,
                      int derivative, double* out) {
  for (int d = 0; d < D; ++d) out[d] = poly_evaluate(N, seg_coeffs + d * N, t, derivative);
}

int traj_evaluate(int N, int D, int K, const double* coeffs, const double* times,
                  double t, int derivative, double* out) {
  double accumulated = 0.0;
  int i = 0;
  for (i = 0; i < K; ++i) {
    accumulated += times[i];
    if (accumulated > t) break;
  }
  if (t > accumulated) {
    for (int d = 0; d < D; ++d) out[d] = 0.0;
    return -1;
  }
  if (i >= K) i = K - 1;
  accumulated -= times[i];
  segment_evaluate(N, D, coeffs + static_cast<size_t>(i) * D * N, t - accumulated, derivative, out);
  return i;
}

double norm_d(const double* v, int D) {
  double s = 0.0;
  for (int d = 0; d < D; ++d) s += v[d] * v[d];
  return std::sqrt(s);
}

// ------------------------------------------------------------- R1 wrapper
int find_roots_jt(const double* inc, int n, double* re, double* im, int* n_roots) {
  *n_roots = 0;
#ifdef MTG_ORACLE_NO_RPOLY
  (void)inc; (void)n; (void)re; (void)im;
  return -2;
#else
  // RPOLY_C:57-68 strip high-order zeros (|c| >= DBL_MIN counts as non-zero)
  int last = -1;
  for (int i = n - 1; i != -1; --i)
    if (std::fabs(inc[i]) >= std::numeric_limits<double>::min()) { last = i; break; }
  if (last == -1) return 1;      // all-zero polynomial: no roots, success
  const int n_coeff = last + 1;
  if (n_coeff < 2) return 1;     // constant: no roots, success
  int degree = n_coeff - 1;
  double poly[101], rr[100], ri[100];
  for (int i = 0; i < n_coeff; ++i) poly[i] = inc[last - i];  // decreasing powers
  mtg_ref_rpoly(poly, &degree, rr, ri);
  if (degree > 0) {
    *n_roots = degree;
    for (int i = 0; i < degree; ++i) { re[i] = rr[i]; im[i] = ri[i]; }
    return 1;
  }
  return 0;
#endif
}

// POLY_C:32-63
int select_candidates_from_roots(double t_start, double t_end, const double* re,
                                 const double* im, int n_roots, double* cand) {
  if (t_start > t_end) return -1;
  int n = 0;
  cand[n++] = t_start;
  cand[n++] = t_end;
  for (int i = 0; i < n_roots; ++i) {
    if (std::fabs(im[i]) > std::numeric_limits<double>::epsilon()) continue;
    const double c = re[i];
    if (c < t_start || c > t_end) continue;
    cand[n++] = c;
  }
  return n;
}

void derivative_coefficients(int N, const double* c, int derivative, double* out) {
  if (derivative == 0) {
    for (int j = 0; j < N; ++j) out[j] = c[j];
    return;
  }
  for (int j = 0; j < N; ++j) out[j] = 0.0;
  for (int j = 0; j < N - derivative; ++j) out[j] = c[j + derivative] * Bc(derivative, j + derivative);
}

void convolve(const double* data, int nd, const double* kernel, int nk, double* out) {
  const int len = nd + nk - 1;
  for (int i = 0; i < len; ++i) {
    out[i] = 0.0;
    const int data_idx = i - nk + 1;
    const int lower = std::max(0, -data_idx);
    const int upper = std::min(nk, nd - data_idx);
    for (int k = lower; k < upper; ++k) out[i] += kernel[nk - 1 - k] * data[data_idx + k];
  }
}

// POLY_C:65-83 on a polynomial with n_c coefficients: roots of derivative+1.
int poly_min_max_candidates(int n_c, const double* c, double t_start, double t_end,
                            int derivative, double* cand) {
  if (n_c - derivative - 1 < 0) return -1;
  double dc[MTGO_MAX_CONV + 2];
  derivative_coefficients(n_c, c, derivative + 1, dc);
  double re[100], im[100];
  int n_roots = 0;
  find_roots_jt(dc, n_c, re, im, &n_roots);  // failure only logged (POLY_C:75-78)
  return select_candidates_from_roots(t_start, t_end, re, im, n_roots, cand);
}

// SEG_C:82-133
int segment_candidate_times(int N, int D, const double* seg, int derivative,
                            double t_start, double t_end, const int* dims,
                            int n_dims, double* cand) {
  if (n_dims <= 0) return -1;
  if (n_dims > 1) {
    const int n_d = N - derivative, n_dd = n_d - 1;
    if (n_dd < 1) return -1;
    const int len = n_d + n_dd - 1;
    double conv[MTGO_MAX_CONV + 2], tmp[MTGO_MAX_CONV + 2], d[MTGO_MAX_N], dd[MTGO_MAX_N];
    for (int i = 0; i < len; ++i) conv[i] = 0.0;
    for (int q = 0; q < n_dims; ++q) {
      const int dim = dims[q];
      if (dim < 0 || dim >= D) return -1;
      derivative_coefficients(N, seg + dim * N, derivative, d);
      derivative_coefficients(N, seg + dim * N, derivative + 1, dd);
      convolve(d, n_d, dd, n_dd, tmp);
      for (int i = 0; i < len; ++i) conv[i] += tmp[i];
    }
    return poly_min_max_candidates(len, conv, t_start, t_end, -1, cand);
  }
  return poly_min_max_candidates(N, seg + dims[0] * N, t_start, t_end, derivative, cand);
}

double segment_magnitude(int N, const double* seg, double t, int derivative,
                         const int* dims, int n_dims) {
  double m = 0.0;
  for (int q = 0; q < n_dims; ++q) {
    const double v = poly_evaluate(N, seg + dims[q] * N, t, derivative);
    m += std::pow(v, 2);
  }
  return std::sqrt(m);
}

// ------------------------------------------------------------- T1 geometry
struct TubeSeg {
  double A[9], b[3], n[3], p_start[3], p_end[3], r_tube, r_sphere, pad[1];
};
static_assert(sizeof(TubeSeg) == 24 * sizeof(double), "geom layout");

void tube_geometry(int K, const double* pos, const double* radii, TubeSeg* g) {
  for (int i = 0; i < K; ++i) {
    const double* s = pos + 3 * i;
    const double* e = pos + 3 * (i + 1);
    double v[3] = {e[0] - s[0], e[1] - s[1], e[2] - s[2]};
    const double nrm = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int k = 0; k < 3; ++k) v[k] = v[k] / nrm;
    const double nx = v[0], ny = v[1], nz = v[2];
    const double px = s[0], py = s[1], pz = s[2];
    TubeSeg& t = g[i];
    const double A[9] = {1 - std::pow(nx, 2), -nx * ny, -nx * nz,
                         -nx * ny, 1 - std::pow(ny, 2), -ny * nz,
                         -nx * nz, -ny * nz, 1 - std::pow(nz, 2)};
    for (int k = 0; k < 9; ++k) t.A[k] = (A[k] > -0.000001 && A[k] < 0.000001) ? 0.0 : A[k];
    const double b[3] = {(std::pow(nx, 2) - 1) * px + nx * ny * py + nx * nz * pz,
                         nx * ny * px + (std::pow(ny, 2) - 1) * py + ny * nz * pz,
                         nx * nz * px + ny * nz * py + (std::pow(nz, 2) - 1) * pz};
    for (int k = 0; k < 3; ++k) t.b[k] = (b[k] > -0.000001 && b[k] < 0.000001) ? 0.0 : b[k];
    const double r_start = (i == 0) ? radii[0] : radii[2 * (i - 1) + 1];
    const double r_end = radii[2 * i + 1];
    for (int k = 0; k < 3; ++k) {
      t.n[k] = v[k];
      t.p_start[k] = s[k] + (-v[k]) * r_start;
      t.p_end[k] = e[k] + v[k] * r_end;
    }
    t.r_tube = radii[2 * i];
    t.r_sphere = radii[2 * i + 1];
    t.pad[0] = 0.0;
  }
}

int tube_flags(const TubeSeg& t, const double* vertex_end, const double* x) {
  double q = 0.0;
  for (int r = 0; r < 3; ++r) {
    const double y = t.A[3 * r] * x[0] + t.A[3 * r + 1] * x[1] + t.A[3 * r + 2] * x[2] + t.b[r];
    q += y * y;
  }
  const bool in_cyl = (q - t.r_tube * t.r_tube) <= 0.0;
  double cs = 0.0, ce = 0.0;
  for (int k = 0; k < 3; ++k) {
    cs += (-t.n[k]) * (x[k] - t.p_start[k]);
    ce += t.n[k] * (x[k] - t.p_end[k]);
  }
  const bool in_caps = (cs <= 0.0) && (ce <= 0.0);
  double sq = 0.0;
  for (int k = 0; k < 3; ++k) sq += (x[k] - vertex_end[k]) * (x[k] - vertex_end[k]);
  const bool in_sphere = (sq - t.r_sphere * t.r_sphere) <= 0.0;
  return (in_cyl && in_caps ? 1 : 0) | (in_sphere ? 2 : 0);
}


// ------------------------------------------------------------- N2  QC_I:267-474
// setupInverseControlPointMappingMatrix QC_I:267-319: control_point_mapping_coefficients(l, j) =
// n!/(n-l)! (-1)^(l+j) / T^l binom(l, j) (j <= l, n = N - 1), inverted numerically, entries in
// (-1e-5, 1e-5) zeroed (:300-306), B_lr_inv = rows reversed, column i times (-1)^i (:308-313).
double factorial_d(int n) {
  double r = 1.0;
  for (int i = 2; i <= n; ++i) r *= i;
  return r;
}
double binomial_d(int n, int k) { return factorial_d(n) / (factorial_d(k) * factorial_d(n - k)); }

bool inverse_control_point_mapping(int N, double T, double* B_inv) {
  const int h = N / 2, n = N - 1;
  Vec M(static_cast<size_t>(h) * h, 0.0), Minv(static_cast<size_t>(h) * h, 0.0);
  M[0] = 1.0;
  for (int l = 1; l < h; ++l)
    for (int j = 0; j < h; ++j)
      if (j <= l)
        M[l * h + j] = factorial_d(n) / factorial_d(n - l) * std::pow(-1.0, l + j) / std::pow(T, l) * binomial_d(l, j);
  if (!general_inverse(h, M.data(), Minv.data())) return false;
  for (int k = 0; k < h; ++k)
    for (int i = 0; i < h; ++i)
      if (Minv[k * h + i] > -0.00001 && Minv[k * h + i] < 0.00001) Minv[k * h + i] = 0.0;
  for (int i = 0; i < N * N; ++i) B_inv[i] = 0.0;
  for (int k = 0; k < h; ++k)
    for (int i = 0; i < h; ++i) {
      B_inv[k * N + i] = Minv[k * h + i];                                              // top-left
      B_inv[(h + k) * N + (h + i)] = Minv[(h - 1 - k) * h + i] * std::pow(-1.0, i);    // bottom-right
    }
  return true;
}
}  // namespace

// =========================================================== C entry points
extern "C" {

void mtgo_base_coefficients(double* out) {
  for (int n = 0; n < MTGO_MAX_CONV; ++n)
    for (int i = 0; i < MTGO_MAX_CONV; ++i) out[n * MTGO_MAX_CONV + i] = kBase.v[n][i];
}
void mtgo_quadratic_cost_jacobian(int N, int derivative, double t, double* Q) {
  quadratic_cost_jacobian(N, derivative, t, Q);
}
void mtgo_base_coeffs_with_time(int N, int derivative, double t, double* out) {
  base_coeffs_with_time(N, derivative, t, out);
}
void mtgo_setup_mapping_matrix(int N, double t, double* A) { setup_mapping_matrix(N, t, A); }
void mtgo_invert_mapping_matrix(int N, const double* A, double* Ainv) {
  invert_mapping_matrix(N, A, Ainv);
}
int mtgo_general_inverse(int n, const double* M, double* Minv) {
  return general_inverse(n, M, Minv) ? 0 : -1;
}

int mtgo_create_random_vertices(int maximum_derivative, int n_segments, int D,
                                const double* pos_min, const double* pos_max,
                                uint64_t seed, int half_n, uint8_t* mask,
                                double* values) {
  if (n_segments < 1 || maximum_derivative <= 0) return -1;  // VTX_C:31-34
  std::mt19937 generator(seed);
  std::vector<std::uniform_real_distribution<double> > dist(D);
  for (int i = 0; i < D; ++i) dist[i] = std::uniform_real_distribution<double>(pos_min[i], pos_max[i]);
  const double min_distance = 0.2;
  const int n_vertices = n_segments + 1;
  std::memset(mask, 0, static_cast<size_t>(n_vertices) * half_n);
  for (size_t i = 0; i < static_cast<size_t>(n_vertices) * half_n * D; ++i) values[i] = 0.0;
  Vec last(D), pos(D);
  for (int i = 0; i < D; ++i) last[i] = dist[i](generator);
  auto set_pos = [&](int v, const Vec& p) {
    mask[v * half_n + 0] = 1;
    for (int d = 0; d < D; ++d) values[(static_cast<size_t>(v) * half_n + 0) * D + d] = p[d];
  };
  auto make_start_or_end = [&](int v, const Vec& p) {  // VTX_C:147-153
    set_pos(v, p);
    for (int k = 1; k <= maximum_derivative && k < half_n; ++k) mask[v * half_n + k] = 1;
  };
  make_start_or_end(0, last);
  for (int v = 1; v < n_vertices; ++v) {
    while (true) {
      for (int d = 0; d < D; ++d) pos[d] = dist[d](generator);
      double s = 0.0;
      for (int d = 0; d < D; ++d) s += (pos[d] - last[d]) * (pos[d] - last[d]);
      if (std::sqrt(s) > min_distance) break;
    }
    set_pos(v, pos);
    last = pos;
  }
  make_start_or_end(n_vertices - 1, last);
  return n_vertices;
}

void mtgo_estimate_segment_times_nfabian(int K, int D, const double* positions,
                                         double v_max, double a_max,
                                         double magic, double* times) {
  for (int i = 0; i < K; ++i) {
    double s = 0.0;
    for (int d = 0; d < D; ++d) {
      const double e = positions[(i + 1) * D + d] - positions[i * D + d];
      s += e * e;
    }
    const double distance = std::sqrt(s);
    times[i] = distance / v_max * 2 *
               (1.0 + magic * v_max / a_max * std::exp(-distance / v_max * 2));
  }
}

void mtgo_estimate_segment_times_velocity_ramp(int K, int D, const double* positions,
                                               double v_max, double a_max,
                                               double* times) {
  for (int i = 0; i < K; ++i) {
    double s = 0.0;
    for (int d = 0; d < D; ++d) {
      const double e = positions[i * D + d] - positions[(i + 1) * D + d];
      s += e * e;
    }
    const double distance = std::sqrt(s);
    const double acc_time = v_max / a_max;
    const double acc_distance = 0.5 * v_max * acc_time;
    times[i] = (distance < 2.0 * acc_distance)
                   ? 2.0 * std::sqrt(distance / a_max)
                   : 2.0 * acc_time + (distance - 2.0 * acc_distance) / v_max;
  }
}

int mtgo_solve(int N, int D, int K, int derivative, const double* times,
               const uint8_t* mask, const double* values, int solver,
               double* coeffs, double* cost, int* counts, double* d_f,
               double* d_p, double* R_out, int* col_of_row) {
  Problem p;
  int rc = setup_problem(p, N, D, K, derivative, times, mask, values);
  if (rc) return rc;
  Vec R;
  rc = solve_linear(p, solver, coeffs, R_out ? &R : nullptr);
  if (rc) return rc;
  if (cost) *cost = compute_cost(p, coeffs);
  if (counts) { counts[0] = p.n_all; counts[1] = p.n_fixed; counts[2] = p.n_free; }
  if (d_f) std::copy(p.d_f.begin(), p.d_f.end(), d_f);
  if (d_p) std::copy(p.d_p.begin(), p.d_p.end(), d_p);
  if (R_out) std::copy(R.begin(), R.end(), R_out);
  if (col_of_row) std::copy(p.col_of_row.begin(), p.col_of_row.end(), col_of_row);
  return 0;
}

int mtgo_solve_canonical_batch(int B, int N, int D, int K, int derivative,
                               const double* positions, const double* times,
                               int solver, int n_threads, double* coeffs,
                               double* cost) {
  const int h = N / 2;
  int bad = 0;
#ifdef _OPENMP
#pragma omp parallel for num_threads(n_threads > 0 ? n_threads : 1) schedule(static) reduction(+ : bad)
#else
  (void)n_threads;
#endif
  for (int b = 0; b < B; ++b) {
    std::vector<uint8_t> mask(static_cast<size_t>(K + 1) * h, 0);
    Vec values(static_cast<size_t>(K + 1) * h * D, 0.0);
    for (int v = 0; v <= K; ++v) {
      mask[v * h] = 1;
      for (int d = 0; d < D; ++d)
        values[(static_cast<size_t>(v) * h) * D + d] = positions[(static_cast<size_t>(b) * (K + 1) + v) * D + d];
    }
    for (int k = 1; k < h; ++k) mask[k] = mask[K * h + k] = 1;
    double c = 0.0;
    const int rc = mtgo_solve(N, D, K, derivative, times + static_cast<size_t>(b) * K, mask.data(),
                              values.data(), solver, coeffs + static_cast<size_t>(b) * K * D * N, &c,
                              nullptr, nullptr, nullptr, nullptr, nullptr);
    if (rc) ++bad;
    if (cost) cost[b] = c;
  }
  return bad ? -1 : 0;
}

int mtgo_coeffs_from_free_constraints(int N, int D, int K, const double* times,
                                      const uint8_t* mask, const double* values,
                                      const double* d_p, double* coeffs) {
  Problem p;
  const int rc = setup_problem(p, N, D, K, N / 2 - 1, times, mask, values);
  if (rc) return rc;
  p.d_p.assign(d_p, d_p + static_cast<size_t>(D) * p.n_free);
  update_segments_from_compact(p, coeffs);
  return 0;
}

int mtgo_cost_time_fd(int N, int D, int K, int derivative, const double* times,
                      const uint8_t* mask, const double* values,
                      const double* d_p, double increment_time, int central,
                      double* J_nominal, double* J_plus, double* J_minus,
                      double* grad_d) {
  Problem p;
  int rc = setup_problem(p, N, D, K, derivative, times, mask, values);
  if (rc) return rc;
  p.d_p.assign(d_p, d_p + static_cast<size_t>(D) * p.n_free);
  const double J0 = cost_derivative(p);
  if (J_nominal) *J_nominal = J0;
  Vec t(K);
  for (int n = 0; n < K; ++n) {
    double J_small = 0.0;
    if (central) {
      for (int i = 0; i < K; ++i) t[i] = times[i];
      t[n] = t[n] <= 0.1 ? 0.1 : t[n] - increment_time;
      rc = update_segment_times(p, t.data());
      if (rc) return rc;
      J_small = cost_derivative(p);
      if (J_minus) J_minus[n] = J_small;
    }
    for (int i = 0; i < K; ++i) t[i] = times[i];
    t[n] = t[n] <= 0.1 ? 0.1 : t[n] + increment_time;
    rc = update_segment_times(p, t.data());
    if (rc) return rc;
    const double J_big = cost_derivative(p);
    if (J_plus) J_plus[n] = J_big;
    if (grad_d)
      grad_d[n] = central ? (J_big - J_small) / (2.0 * increment_time)
                          : (J_big - J0) / (increment_time);
  }
  return 0;
}

double mtgo_poly_evaluate(int N, const double* c, double t, int derivative) {
  return poly_evaluate(N, c, t, derivative);
}
void mtgo_poly_derivative_coefficients(int N, const double* c, int derivative, double* out) {
  derivative_coefficients(N, c, derivative, out);
}
void mtgo_convolve(const double* data, int n_data, const double* kernel, int n_kernel, double* out) {
  convolve(data, n_data, kernel, n_kernel, out);
}
int mtgo_traj_evaluate(int N, int D, int K, const double* coeffs, const double* times,
                       double t, int derivative, double* out) {
  return traj_evaluate(N, D, K, coeffs, times, t, derivative, out);
}

int mtgo_traj_evaluate_range(int N, int D, int K, const double* coeffs,
                             const double* times, double t_start, double t_end,
                             double dt, int derivative, int cap, double* out,
                             double* sampling_times, int32_t* segment_idx) {
  double accumulated = 0.0;
  int i = 0;
  for (i = 0; i < K; ++i) {
    accumulated += times[i];
    if (accumulated > t_start) break;
  }
  if (t_start > accumulated) return -1;
  if (i >= K) return -1;  // t_start == max_time: reference indexes segments_[K] (UB)
  accumulated -= times[i];
  double tau = t_start - accumulated;
  int n = 0;
  while (accumulated < t_end) {
    if (tau > times[i]) {
      tau = tau - times[i];
      ++i;
      if (i >= K) break;
      continue;
    }
    if (n >= cap) return n;
    if (out) segment_evaluate(N, D, coeffs + static_cast<size_t>(i) * D * N, tau, derivative, out + static_cast<size_t>(n) * D);
    if (sampling_times) sampling_times[n] = accumulated;
    if (segment_idx) segment_idx[n] = i;
    ++n;
    tau += dt;
    accumulated += dt;
  }
  return n;
}

int mtgo_find_roots_jenkins_traub(const double* inc, int n, double* re, double* im, int* n_roots) {
  return find_roots_jt(inc, n, re, im, n_roots);
}
int mtgo_select_min_max_candidates_from_roots(double t_start, double t_end, const double* re,
                                              const double* im, int n_roots, double* cand) {
  return select_candidates_from_roots(t_start, t_end, re, im, n_roots, cand);
}

int mtgo_poly_compute_min_max(int N, const double* c, double t_start, double t_end,
                              int derivative, double* min_t, double* min_v,
                              double* max_t, double* max_v) {
  double cand[128];
  const int n = poly_min_max_candidates(N, c, t_start, t_end, derivative, cand);
  if (n <= 0) return -1;
  *min_t = cand[0]; *min_v = std::numeric_limits<double>::max();
  *max_t = cand[0]; *max_v = std::numeric_limits<double>::lowest();
  for (int i = 0; i < n; ++i) {  // POLY_C:116-143
    const double v = poly_evaluate(N, c, cand[i], derivative);
    if (v < *min_v) { *min_t = cand[i]; *min_v = v; }
    if (v > *max_v) { *max_t = cand[i]; *max_v = v; }
  }
  return 0;
}

int mtgo_segment_candidate_times(int N, int D, const double* seg, int derivative,
                                 double t_start, double t_end, const int* dims,
                                 int n_dims, double* cand, int cap) {
  double tmp[128];
  const int n = segment_candidate_times(N, D, seg, derivative, t_start, t_end, dims, n_dims, tmp);
  for (int i = 0; i < n && i < cap; ++i) cand[i] = tmp[i];
  return n;
}

int mtgo_traj_min_max_magnitude(int N, int D, int K, const double* coeffs,
                                const double* times, int derivative,
                                const int* dims, int n_dims, double* min_time,
                                double* min_value, int* min_seg,
                                double* max_time, double* max_value, int* max_seg) {
  double mn_v = std::numeric_limits<double>::max(), mx_v = std::numeric_limits<double>::lowest();
  double mn_t = 0.0, mx_t = 0.0;
  int mn_s = 0, mx_s = 0;
  for (int s = 0; s < K; ++s) {
    const double* seg = coeffs + static_cast<size_t>(s) * D * N;
    double cand[128];
    // SEG_C:135-158: the bool of the candidate-time search is ignored there.
    int n = segment_candidate_times(N, D, seg, derivative, 0.0, times[s], dims, n_dims, cand);
    if (n < 0) n = 0;
    // SEG_C:160-184
    if (0.0 > times[s]) return -1;
    double smn_v = std::numeric_limits<double>::max(), smx_v = std::numeric_limits<double>::lowest();
    double smn_t = 0.0, smx_t = 0.0;
    for (int i = 0; i < n; ++i) {
      const double t = cand[i];
      if (t < 0.0 || t > times[s]) continue;
      const double m = segment_magnitude(N, seg, t, derivative, dims, n_dims);
      if (smx_v < m) { smx_v = m; smx_t = t; }   // std::max keeps first on ties
      if (m < smn_v) { smn_v = m; smn_t = t; }   // std::min keeps first on ties
    }
    if (smn_v < mn_v) { mn_v = smn_v; mn_t = smn_t; mn_s = s; }
    if (smx_v > mx_v) { mx_v = smx_v; mx_t = smx_t; mx_s = s; }
  }
  *min_time = mn_t; *min_value = mn_v; *min_seg = mn_s;
  *max_time = mx_t; *max_value = mx_v; *max_seg = mx_s;
  return 0;
}

int mtgo_opt_max_magnitude(int N, int D, int K, const double* coeffs,
                           const double* times, int derivative,
                           double* max_time, double* max_value, int* max_seg) {
  if (!(N - derivative - 1 > 0)) return -1;  // LIN_I:401
  std::vector<int> dims(D);
  for (int d = 0; d < D; ++d) dims[d] = d;
  double e_t = 0.0, e_v = 0.0;  // Extremum() default
  int e_s = 0;
  Vec val(D);
  for (int s = 0; s < K; ++s) {
    const double* seg = coeffs + static_cast<size_t>(s) * D * N;
    double cand[128];
    int n = segment_candidate_times(N, D, seg, derivative, 0.0, times[s], dims.data(), D, cand);
    if (n < 0) n = 0;  // the pushed 0.0 is cleared by the callee (SEG_C:87)
    for (int i = 0; i < n; ++i) {
      segment_evaluate(N, D, seg, cand[i], derivative, val.data());
      const double m = norm_d(val.data(), D);
      if (e_v < m) { e_v = m; e_t = cand[i]; e_s = s; }
    }
  }
  const double* seg = coeffs + static_cast<size_t>(K - 1) * D * N;
  segment_evaluate(N, D, seg, times[K - 1], derivative, val.data());
  const double m = norm_d(val.data(), D);
  if (e_v < m) { e_v = m; e_t = times[K - 1]; e_s = K - 1; }
  *max_time = e_t; *max_value = e_v; *max_seg = e_s;
  return 0;
}

double mtgo_sampled_maximum_magnitude(int N, int D, int K, const double* coeffs,
                                      const double* times, int derivative, double dt) {
  double max_time = 0.0;
  for (int i = 0; i < K; ++i) max_time += times[i];
  double maximum = -1e9;
  Vec v(D);
  for (double ts = 0; ts < max_time; ts += dt) {
    traj_evaluate(N, D, K, coeffs, times, ts, derivative, v.data());
    const double cur = norm_d(v.data(), D);
    if (cur > maximum) maximum = cur;
  }
  return maximum;
}

double mtgo_cost_numeric(int N, int D, int K, const double* coeffs,
                         const double* times, int derivative, double dt) {
  double max_time = 0.0;
  for (int i = 0; i < K; ++i) max_time += times[i];
  double cost = 0.0;
  Vec v(D);
  for (double ts = 0; ts < max_time; ts += dt) {
    traj_evaluate(N, D, K, coeffs, times, ts, derivative, v.data());
    double s = 0.0;
    for (int d = 0; d < D; ++d) s += v[d] * v[d];
    cost += s * dt;
  }
  return cost;
}

void mtgo_tube_geometry(int K, const double* positions, const double* radii, double* geom) {
  tube_geometry(K, positions, radii, reinterpret_cast<TubeSeg*>(geom));
}
int mtgo_tube_flags(const double* geom_seg, const double* vertex_end, const double* x) {
  return tube_flags(*reinterpret_cast<const TubeSeg*>(geom_seg), vertex_end, x);
}

int mtgo_feasibility_sweep(int N, int K, const double* coeffs, const double* times,
                           const double* positions, const double* radii,
                           double v_max, double a_max, double t_start,
                           double t_end, double dt, int cap, double* pos_out,
                           uint8_t* flags, double* max_v, double* max_a) {
  const int D = 3;
  std::vector<TubeSeg> geom;
  if (radii) { geom.resize(K); tube_geometry(K, positions, radii, geom.data()); }
  double accumulated = 0.0;
  int i = 0;
  for (i = 0; i < K; ++i) {
    accumulated += times[i];
    if (accumulated > t_start) break;
  }
  if (t_start > accumulated || i >= K) return -1;
  accumulated -= times[i];
  double tau = t_start - accumulated;
  int n = 0;
  double mv = 0.0, ma = 0.0;
  while (accumulated < t_end) {
    if (tau > times[i]) {
      tau = tau - times[i];
      ++i;
      if (i >= K) break;
      continue;
    }
    if (n >= cap) break;
    const double* seg = coeffs + static_cast<size_t>(i) * D * N;
    double x[3], v[3], a[3];
    segment_evaluate(N, D, seg, tau, 0, x);
    segment_evaluate(N, D, seg, tau, 1, v);
    segment_evaluate(N, D, seg, tau, 2, a);
    const double nv = norm_d(v, 3), na = norm_d(a, 3);
    if (nv > mv) mv = nv;
    if (na > ma) ma = na;
    int f = (nv <= v_max ? 1 : 0) | (na <= a_max ? 2 : 0);
    if (radii) f |= (tube_flags(geom[i], positions + 3 * (i + 1), x) & 1) ? 4 : 0;
    if (pos_out) for (int d = 0; d < 3; ++d) pos_out[static_cast<size_t>(n) * 3 + d] = x[d];
    if (flags) flags[n] = static_cast<uint8_t>(f);
    ++n;
    tau += dt;
    accumulated += dt;
  }
  if (max_v) *max_v = mv;
  if (max_a) *max_a = ma;
  return n;
}


/* N2 */
int mtgo_inverse_control_point_mapping(int N, double T, double* B_inv) {
  return inverse_control_point_mapping(N, T, B_inv) ? 0 : -1;
}

// Control points of every segment (F B_inv C [d_f; d_p], QC_I:321-355: control point j of segment i =
// row j of B_inv_i times the segment's endpoint derivatives) and the values of the reference's
// constraints on them (feasible <=> value <= 0): tube x^T LL x + L x + mu on control points 1..N-2
// (QC_I:369-429), the two end-cap half spaces on the same points (:431-474), the sphere on the last
// control point of every segment but the last (:357-365, :349-351).
int mtgo_control_point_constraints(int N, int K, int D, const double* derivatives, const double* times,
                                   const double* positions, const double* radii, double* control_points,
                                   double* tube, double* cap_start, double* cap_end, double* sphere) {
  const int h = N / 2;
  std::vector<TubeSeg> geom;
  const bool constraints = positions && radii && D == 3;
  if (constraints) { geom.resize(K); tube_geometry(K, positions, radii, geom.data()); }
  Vec B_inv(static_cast<size_t>(N) * N), cp(static_cast<size_t>(N) * D);
  for (int i = 0; i < K; ++i) {
    if (!inverse_control_point_mapping(N, times[i], B_inv.data())) return -1;
    for (int dim = 0; dim < D; ++dim)
      for (int j = 0; j < N; ++j) {
        double s = 0.0;
        for (int c = 0; c < N; ++c) {
          const int v = i + (c >= h ? 1 : 0), k = c % h;
          s += B_inv[j * N + c] * derivatives[(static_cast<size_t>(v) * h + k) * D + dim];
        }
        cp[static_cast<size_t>(j) * D + dim] = s;
        if (control_points) control_points[(static_cast<size_t>(i) * N + j) * D + dim] = s;
      }
    if (!constraints) continue;
    const TubeSeg& t = geom[i];
    double LL[9], L[3], mu = 0.0;  // LL = A^T A, L = 2 b^T A, mu = b^T b - r^2   (QC_I:412-415)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += t.A[3 * k + r] * t.A[3 * k + c];
        LL[3 * r + c] = s;
      }
    for (int c = 0; c < 3; ++c) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += t.b[k] * t.A[3 * k + c];
      L[c] = 2 * s;
    }
    for (int k = 0; k < 3; ++k) mu += t.b[k] * t.b[k];
    mu -= std::pow(t.r_tube, 2);
    for (int j = 1; j < N - 1; ++j) {
      const double* x = &cp[static_cast<size_t>(j) * 3];
      double q = 0.0, lin = 0.0;
      for (int r = 0; r < 3; ++r) {
        double s = 0.0;
        for (int c = 0; c < 3; ++c) s += LL[3 * r + c] * x[c];
        q += x[r] * s;
        lin += L[r] * x[r];
      }
      if (tube) tube[static_cast<size_t>(i) * (N - 2) + (j - 1)] = q + lin + mu;
      double cs = 0.0, ce = 0.0;  // norm_vec_start . x - norm_vec_start . p_start ; norm_vec_end . x - norm_vec_end . p_end
      for (int k = 0; k < 3; ++k) {
        cs += (-t.n[k]) * x[k] - (-t.n[k]) * t.p_start[k];
        ce += t.n[k] * x[k] - t.n[k] * t.p_end[k];
      }
      if (cap_start) cap_start[static_cast<size_t>(i) * (N - 2) + (j - 1)] = cs;
      if (cap_end) cap_end[static_cast<size_t>(i) * (N - 2) + (j - 1)] = ce;
    }
    if (sphere) {
      if (i < K - 1) {
        const double* x = &cp[static_cast<size_t>(N - 1) * 3];
        const double* pv = positions + 3 * (i + 1);
        double xx = 0.0, px = 0.0, pp = 0.0;  // x^T x - 2 p^T x + p^T p - r^2   (QC_I:357-365)
        for (int k = 0; k < 3; ++k) { xx += x[k] * x[k]; px += pv[k] * x[k]; pp += pv[k] * pv[k]; }
        sphere[i] = xx - 2 * px + pp - std::pow(t.r_sphere, 2);
      } else {
        sphere[i] = -std::numeric_limits<double>::infinity();
      }
    }
  }
  return 0;
}

// N3  NL_I:2907-3003 printMatlabSampledTrajectory as a matrix: rows [t, pos(D), vel(D), acc(D), jerk(D),
// snap(D), tm]; per segment `for (t = 0; t < T_i; t += dt)`, values as T_seg^T (derivative matrix) p with
// T_seg[n] = pow(t, n) and the super-diagonal derivative matrices of NL_I:2820-2846; rows without a sample
// stay zero; output(i, 1 + 5 D) = end time of segment i. Returns the number of sample rows (j).
int mtgo_sample_dump(int N, int D, int K, const double* coeffs, const double* times, double dt, int max_rows,
                     double* rows) {
  const int W = 5 * D + 2;
  for (size_t e = 0; e < static_cast<size_t>(max_rows) * W; ++e) rows[e] = 0.0;
  int j = 0;
  double current_segment_time = 0.0;
  for (int i = 0; i < K; ++i) {
    for (double t = 0.0; t < times[i]; t += dt) {
      if (j < max_rows) {
        double* r = rows + static_cast<size_t>(j) * W;
        r[0] = t + current_segment_time;
        for (int k = 0; k < D; ++k) {
          const double* p = coeffs + (static_cast<size_t>(i) * D + k) * N;
          for (int der = 0; der < 5; ++der) {
            double v = 0.0;  // sum_n pow(t, n) * (n+1)...(n+der) * p[n + der]
            for (int n = 0; n + der < N; ++n) {
              double f = 1.0;
              for (int q = 1; q <= der; ++q) f *= (n + q);
              v += std::pow(t, n) * (f * p[n + der]);
            }
            r[1 + der * D + k] = v;
          }
        }
        ++j;
      }
    }
    current_segment_time += times[i];
    if (i < max_rows) rows[static_cast<size_t>(i) * W + (W - 1)] = current_segment_time;
  }
  return j;
}

// N1  NL_I:1537-1606 getCostAndGradientDerivative: J_d and grad_{d_p} = 2 R_pf d_f + 2 R_pp d_p from the dense R
int mtgo_cost_gradient_derivative(int N, int D, int K, int derivative, const double* times, const uint8_t* mask,
                                  const double* values, const double* d_p, double* J_d, double* grad) {
  Problem p;
  int rc = setup_problem(p, N, D, K, derivative, times, mask, values);
  if (rc) return rc;
  p.d_p.assign(d_p, d_p + static_cast<size_t>(D) * p.n_free);
  if (J_d) *J_d = cost_derivative(p);
  Vec R;
  construct_R(p, &R);
  const int nf = p.n_fixed, np = p.n_free, n = nf + np;
  for (int dim = 0; dim < D; ++dim)
    for (int r = 0; r < np; ++r) {
      double s1 = 0.0, s2 = 0.0;  // (2 d_f^T R_pf^T)_r + (2 d_p^T R_pp)_r
      for (int c = 0; c < nf; ++c) s1 += p.d_f[static_cast<size_t>(dim) * nf + c] * R[static_cast<size_t>(nf + r) * n + c];
      for (int c = 0; c < np; ++c) s2 += p.d_p[static_cast<size_t>(dim) * np + c] * R[static_cast<size_t>(nf + c) * n + nf + r];
      grad[static_cast<size_t>(dim) * np + r] = 2 * s1 + 2 * s2;
    }
  return 0;
}

// N1  NL_I:2735-2766 + 2365-2490: the soft-constraint cost of the trajectory with free derivatives d_p and its
// finite-difference gradient (central: (cost(d_p + e) - cost(d_p - e)) / (2 inc); forward: (cost(d_p + e) - J_sc) / inc),
// each evaluation = setFreeConstraints + computeMaximumOfMagnitude (LIN_I:455-487) per constraint.
double soft_constraint_cost(Problem& p, const double* d_p, int n_con, const int* ders, const double* limits,
                            double weight, double max_cost, Vec& coeffs) {
  p.d_p.assign(d_p, d_p + static_cast<size_t>(p.D) * p.n_free);
  update_segments_from_compact(p, coeffs.data());
  double cost = 0.0;
  for (int c = 0; c < n_con; ++c) {
    double t = 0.0, v = 0.0;
    int sidx = 0;
    mtgo_opt_max_magnitude(p.N, p.D, p.K, coeffs.data(), p.times, ders[c], &t, &v, &sidx);
    const double rel = (v - limits[c]) / limits[c];
    cost += std::min(max_cost, std::exp(rel * weight));
  }
  return cost;
}

int mtgo_soft_constraint_gradient(int N, int D, int K, int derivative, const double* times, const uint8_t* mask,
                                  const double* values, const double* d_p, int n_con, const int* ders,
                                  const double* limits, double weight, double max_cost, double increment,
                                  int central, double* J_sc, double* grad) {
  Problem p;
  int rc = setup_problem(p, N, D, K, derivative, times, mask, values);
  if (rc) return rc;
  Vec coeffs(static_cast<size_t>(K) * D * N);
  const size_t nv = static_cast<size_t>(D) * p.n_free;
  Vec x(d_p, d_p + nv);
  const double J = soft_constraint_cost(p, x.data(), n_con, ders, limits, weight, max_cost, coeffs);
  if (J_sc) *J_sc = J;
  if (!grad) return 0;
  for (size_t q = 0; q < nv; ++q) {
    const double keep = x[q];
    x[q] = keep + increment;
    const double right = soft_constraint_cost(p, x.data(), n_con, ders, limits, weight, max_cost, coeffs);
    double left = J;
    if (central) {
      x[q] = keep - increment;
      left = soft_constraint_cost(p, x.data(), n_con, ders, limits, weight, max_cost, coeffs);
    }
    x[q] = keep;
    grad[q] = central ? (right - left) / (2.0 * increment) : (right - J) / increment;
  }
  return 0;
}

// N4  NL_I:1608-1780 getCostAndGradientCollision with getCostAndGradientPotentialOctree (:1783-1917) and
// getCostPotential (:2659-2684); the octree distance (:1920-2043) is replaced by a lookup in a dense grid of
// distances [nx][ny][nz] (metres) whose element [0][0][0] is voxel `origin`. Canonical constraint pattern: the free
// derivative (v, k), v = 1..K-1, k = 1..h-1, is row h + k of segment v-1 and row k of segment v of C.
// grad [3][(K-1)(h-1)] (free_constraints order) or NULL. Like the reference, a collision returns J_c = 0 and keeps
// the gradient accumulated so far (its zeroing loop iterates by value, NL_I:1774-1778).
static double potential_cost(double distance, double epsilon, double robot_radius, double multiplier, bool* collision) {
  *collision = false;
  double cost = 0.0;
  distance -= robot_radius;
  if (distance <= 0.0) {
    cost = multiplier * (-distance) + 0.5 * epsilon;
    *collision = true;
  } else if (distance <= epsilon) {
    const double e = distance - epsilon;
    cost = 0.5 * 1.0 / epsilon * e * e;
  }
  return cost;
}

int mtgo_collision_cost(int N, int K, const double* coeffs, const double* times, const double* grid,
                        const int* size, const int* origin, double res, const double* min_bound,
                        const double* max_bound, double dt, double epsilon, double robot_radius, double multiplier,
                        double* J_c, double* grad, int* in_collision, int* n_checks) {
  const int D = 3, h = N / 2, NF = h - 1;
  auto dist_at = [&](int vx, int vy, int vz) {
    const int ix = vx - origin[0], iy = vy - origin[1], iz = vz - origin[2];
    if (ix < 0 || iy < 0 || iz < 0 || ix >= size[0] || iy >= size[1] || iz >= size[2])
      return std::numeric_limits<double>::max();  // no occupied voxel in reach (getDistanceOctree: DBL_MAX * res)
    return grid[(static_cast<size_t>(ix) * size[1] + iy) * size[2] + iz];
  };
  const int n_free = (K - 1) * NF;
  if (grad) for (int e = 0; e < D * n_free; ++e) grad[e] = 0.0;
  double J = 0.0;
  bool collided = false;
  int checks = 0;
  double prev[3] = {0, 0, 0};
  double time_sum = -1, dist_sum = 0.0, t = 0.0;
  Vec A(static_cast<size_t>(N) * N), Ainv(static_cast<size_t>(N) * N), T(N);
  for (int i = 0; i < K; ++i) {
    setup_mapping_matrix(N, times[i], A.data());
    invert_mapping_matrix(N, A.data(), Ainv.data());
    for (t = 0.0; t < times[i]; t += dt) {
      for (int n = 0; n < N; ++n) T[n] = std::pow(t, n);
      double pos[3], vel[3];
      for (int k = 0; k < D; ++k) {
        const double* pk = coeffs + (static_cast<size_t>(i) * D + k) * N;
        pos[k] = 0.0;
        vel[k] = 0.0;
        for (int n = 0; n < N; ++n) pos[k] += T[n] * pk[n];
        for (int n = 0; n + 1 < N; ++n) vel[k] += T[n] * ((n + 1) * pk[n + 1]);
      }
      if (time_sum < 0) {
        time_sum = 0.0;
        for (int k = 0; k < 3; ++k) prev[k] = pos[k];
        continue;
      }
      time_sum += dt;
      dist_sum += std::sqrt((pos[0] - prev[0]) * (pos[0] - prev[0]) + (pos[1] - prev[1]) * (pos[1] - prev[1]) +
                            (pos[2] - prev[2]) * (pos[2] - prev[2]));
      for (int k = 0; k < 3; ++k) prev[k] = pos[k];
      if (dist_sum < res) continue;
      ++checks;
      bool valid = true;
      for (int k = 0; k < 3; ++k)
        if (pos[k] < min_bound[k] + res || pos[k] > max_bound[k] - res) valid = false;
      const int v[3] = {static_cast<int>(pos[0] / res), static_cast<int>(pos[1] / res), static_cast<int>(pos[2] / res)};
      bool hit = false;
      const double c = potential_cost(valid ? dist_at(v[0], v[1], v[2]) : 0.0, epsilon, robot_radius, multiplier, &hit);
      if (hit) {
        collided = true;
        break;
      }
      const double nv = std::sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
      J += c * nv * time_sum;
      if (grad && nv > 1e-6) {
        double gc[3];
        for (int k = 0; k < 3; ++k) {
          bool d1, d2;
          const double left = potential_cost(dist_at(v[0] - (k == 0), v[1] - (k == 1), v[2] - (k == 2)), epsilon,
                                             robot_radius, multiplier, &d1);
          const double right = potential_cost(dist_at(v[0] + (k == 0), v[1] + (k == 1), v[2] + (k == 2)), epsilon,
                                              robot_radius, multiplier, &d2);
          gc[k] = (right - left) / (2.0 * res);
        }
        // T_all_seg^T L_pp and T_all_seg^T V_all L_pp: only the free derivatives of vertices i and i + 1
        for (int side = 0; side < 2; ++side) {
          const int vtx = i + side;
          if (vtx < 1 || vtx > K - 1) continue;
          for (int kk = 1; kk < h; ++kk) {
            const int row = side == 0 ? kk : h + kk;  // local row of this segment
            double TL = 0.0, TVL = 0.0;
            for (int n = 0; n < N; ++n) TL += T[n] * Ainv[n * N + row];
            for (int n = 0; n + 1 < N; ++n) TVL += T[n] * ((n + 1) * Ainv[(n + 1) * N + row]);
            for (int k = 0; k < D; ++k)
              grad[static_cast<size_t>(k) * n_free + (vtx - 1) * NF + (kk - 1)] +=
                  nv * time_sum * gc[k] * TL + time_sum * c * vel[k] / nv * TVL;
          }
        }
      }
      dist_sum = 0.0;
      time_sum = 0.0;
      for (int k = 0; k < 3; ++k) prev[k] = pos[k];
    }
    if (collided) break;
    time_sum += -dt + (times[i] - t);
  }
  *J_c = collided ? 0.0 : J;
  if (in_collision) *in_collision = collided ? 1 : 0;
  if (n_checks) *n_checks = checks;
  return 0;
}

int mtgo_has_reference_rpoly(void) {
#ifdef MTG_ORACLE_NO_RPOLY
  return 0;
#else
  return 1;
#endif
}

}  // extern "C"

// Host-side constant tables for the closed-form time scaling of the min-snap
// blocks (DESIGN.md "time-scaling identity"):
//
//   A(T)   = S^-1 A(1) P            S = diag(T^alpha), alpha = (0..h-1, 0..h-1)
//   A(T)^-1 = P^-1 A(1)^-1 S        P = diag(T^0 .. T^(N-1))
//   Q(T)   = T^(1-2d) P Q(1) P
//   H(T)   = A^-T Q A^-1 = T^(1-2d) S H1 S,   H1 = A(1)^-T Q(1) A(1)^-1
//
// with A = [A(0); A(T)] the endpoint-derivative map of the reference
// (polynomial_optimization_linear_impl.h:101-111, polynomial.h:201-228) and Q
// its cost Hessian (:557-573). A(1) and Q(1) are small exact rationals; they
// are inverted / multiplied here in binary128 and rounded to double ONCE, so the
// device never inverts an ill-conditioned A(T) (cond 1e6..1e9 in fp64).
//
// Compiled by g++ (not nvcc) because of __float128.
#include "tables.h"

namespace mtg {

namespace {
typedef __float128 q128;

q128 falling_factorial(int n, int i) {  // i!/(i-n)!, src/polynomial.cpp:145-161
  if (i < n) return 0;
  q128 r = 1;
  for (int k = i - n + 1; k <= i; ++k) r *= k;
  return r;
}

bool invert(int n, const q128* M, q128* Minv) {
  q128 a[MTG_TAB_LD * MTG_TAB_LD], b[MTG_TAB_LD * MTG_TAB_LD];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      a[i * n + j] = M[i * n + j];
      b[i * n + j] = (i == j) ? 1 : 0;
    }
  for (int c = 0; c < n; ++c) {
    int piv = c;
    q128 best = a[c * n + c] < 0 ? -a[c * n + c] : a[c * n + c];
    for (int r = c + 1; r < n; ++r) {
      const q128 v = a[r * n + c] < 0 ? -a[r * n + c] : a[r * n + c];
      if (v > best) { best = v; piv = r; }
    }
    if (best == 0) return false;
    if (piv != c)
      for (int j = 0; j < n; ++j) {
        q128 t = a[c * n + j]; a[c * n + j] = a[piv * n + j]; a[piv * n + j] = t;
        t = b[c * n + j]; b[c * n + j] = b[piv * n + j]; b[piv * n + j] = t;
      }
    const q128 inv = q128(1) / a[c * n + c];
    for (int j = 0; j < n; ++j) { a[c * n + j] *= inv; b[c * n + j] *= inv; }
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const q128 f = a[r * n + c];
      if (f == 0) continue;
      for (int j = 0; j < n; ++j) { a[r * n + j] -= f * a[c * n + j]; b[r * n + j] -= f * b[c * n + j]; }
    }
  }
  for (int i = 0; i < n * n; ++i) Minv[i] = b[i];
  return true;
}
}  // namespace

bool compute_tables(int N, int derivative, Tables* out) {
  if (N < 2 || N > MTG_TAB_LD || (N & 1) || derivative < 0 || derivative > N / 2 - 1) return false;
  const int h = N / 2, d = derivative;
  q128 A[MTG_TAB_LD * MTG_TAB_LD], Ai[MTG_TAB_LD * MTG_TAB_LD], Q[MTG_TAB_LD * MTG_TAB_LD];
  q128 tmp[MTG_TAB_LD * MTG_TAB_LD], H[MTG_TAB_LD * MTG_TAB_LD];
  for (int i = 0; i < N * N; ++i) A[i] = Q[i] = 0;
  for (int r = 0; r < h; ++r) {
    A[r * N + r] = falling_factorial(r, r);                      // derivative r at t = 0
    for (int j = r; j < N; ++j) A[(r + h) * N + j] = falling_factorial(r, j);  // at t = 1
  }
  for (int a = d; a < N; ++a)
    for (int b = d; b < N; ++b)
      Q[a * N + b] = falling_factorial(d, a) * falling_factorial(d, b) * 2 / q128(a + b - 2 * d + 1);
  if (!invert(N, A, Ai)) return false;
  for (int i = 0; i < N; ++i)           // tmp = Q Ai
    for (int j = 0; j < N; ++j) {
      q128 s = 0;
      for (int k = 0; k < N; ++k) s += Q[i * N + k] * Ai[k * N + j];
      tmp[i * N + j] = s;
    }
  for (int i = 0; i < N; ++i)           // H = Ai^T tmp (upper), mirrored: exactly symmetric
    for (int j = i; j < N; ++j) {
      q128 s = 0;
      for (int k = 0; k < N; ++k) s += Ai[k * N + i] * tmp[k * N + j];
      H[i * N + j] = H[j * N + i] = s;
    }
  // W = L^T Ainv1[d.., :],  Q[d.., d..] = L L^T (Cholesky in binary128)
  const int nq = N - d;
  q128 L[MTG_TAB_LD * MTG_TAB_LD], Wm[MTG_TAB_LD * MTG_TAB_LD];
  for (int i = 0; i < nq * nq; ++i) L[i] = 0;
  for (int j = 0; j < nq; ++j) {
    q128 s = Q[(d + j) * N + (d + j)];
    for (int k = 0; k < j; ++k) s -= L[j * nq + k] * L[j * nq + k];
    if (!(s > 0)) return false;
    // Newton square root in binary128 from a double seed
    q128 r = static_cast<q128>(__builtin_sqrt(static_cast<double>(s)));
    for (int it = 0; it < 4; ++it) r = (r + s / r) / 2;
    L[j * nq + j] = r;
    for (int i = j + 1; i < nq; ++i) {
      q128 t = Q[(d + i) * N + (d + j)];
      for (int k = 0; k < j; ++k) t -= L[i * nq + k] * L[j * nq + k];
      L[i * nq + j] = t / r;
    }
  }
  for (int i = 0; i < nq; ++i)
    for (int m = 0; m < N; ++m) {
      q128 s = 0;
      for (int a = i; a < nq; ++a) s += L[a * nq + i] * Ai[(d + a) * N + m];
      Wm[i * N + m] = s;
    }
  out->N = N;
  out->derivative = d;
  for (int i = 0; i < MTG_TAB_LD * MTG_TAB_LD; ++i) out->H1[i] = out->Ainv1[i] = out->W[i] = out->Lt[i] = 0.0;
  for (int i = 0; i < nq; ++i)
    for (int a = i; a < nq; ++a) out->Lt[i * MTG_TAB_LD + a] = static_cast<double>(L[a * nq + i]);
  for (int i = 0; i < nq; ++i)
    for (int m = 0; m < N; ++m) out->W[i * MTG_TAB_LD + m] = static_cast<double>(Wm[i * N + m]);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      out->H1[i * MTG_TAB_LD + j] = static_cast<double>(H[i * N + j]);
      out->Ainv1[i * MTG_TAB_LD + j] = static_cast<double>(Ai[i * N + j]);
    }
  for (int n = 0; n < MTG_BASE_LD; ++n)
    for (int i = 0; i < MTG_BASE_LD; ++i)
      out->base[n * MTG_BASE_LD + i] = static_cast<double>(falling_factorial(n, i));
  return true;
}

}  // namespace mtg

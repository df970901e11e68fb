// solve_generic — batched closed-form min-derivative solve for an ARBITRARY per-vertex constraint
// pattern shared by the batch: mask[v][k] = 1 fixes derivative k of vertex v (any subset, at any
// vertex), 0 leaves it to the optimiser. This is the general form of
// setupConstraintReorderingMatrix + constructR + solveLinear + updateSegmentsFromCompactConstraints
// + computeCost (LIN_I:171-252, 306-335, 337-379, 254-275, 113-130); solve_canonical is its
// specialisation to the createRandomVertices pattern.
//
// The normal equations R_pp d_p = -R_pf d_f stay block tridiagonal over the vertices, now with
// blocks of size f_v = number of free derivatives of vertex v (0..h):
//   D_v = H_{v-1}[end F_v, end F_v] + H_v[start F_v, start F_v],  U_v = H_v[start F_v, end F_{v+1}],
//   b_v = -(H_{v-1}[end F_v, fixed] d_f + H_v[start F_v, fixed] d_f),   H_i = T_i^(1-2d) S H1 S,
// solved by the block-Thomas recurrences S_v = D_v - U_{v-1}^T G_{v-1}, G_v = S_v^-1 U_v,
// z_v = S_v^-1 (b_v - U_{v-1}^T z_{v-1}), x_v = z_v - G_v x_{v+1} with Cholesky pivots (a
// non-positive pivot — e.g. position free everywhere — sets MTG_ST_NOT_SPD). One thread per
// trajectory; the mask is uniform across the batch, so there is no divergence; (G_v, z_v) are
// parked in a global scratch, slot-major / batch-minor (coalesced).
#ifndef MTG_SOLVE_GENERIC_CUH_
#define MTG_SOLVE_GENERIC_CUH_

#include <stdint.h>

#include "solve_canonical.cuh"

namespace mtg {

struct SolveGenericParams {
  const uint8_t* __restrict__ mask;    // [(K+1)][h], device, shared by the batch
  const double* __restrict__ values;   // elem ((v*h + k)*D + dim), rec (K+1)*h*D: fixed values (others ignored)
  double* __restrict__ free_out;       // elem (dim*n_free + q), rec D*n_free; or nullptr
  double* __restrict__ scratch;        // [(K+1)][h*h + h*D][nb] parked G_v, z_v (chunk-local)
  int n_free;
};

template <int HN, int D, bool AOS>
__global__ void __launch_bounds__(128) solve_generic_kernel(const SolveCanonicalParams p, const SolveGenericParams g) {
  constexpr int SL = HN * HN + HN * D;
  const int local = blockIdx.x * blockDim.x + threadIdx.x;
  if (local >= p.nb) return;
  const int b = p.b0 + local;
  const int K = p.K, d = p.derivative;
  const size_t B = (size_t)p.B, nb = (size_t)p.nb;
  const size_t rec_v = (size_t)(K + 1) * HN * D;
  uint32_t st = 0;

  auto seg_time = [&](int i) {
    double T = p.seg_times[at<AOS>((size_t)i, (size_t)K, B, b)];
    if (!(T > 0.0) || !(T < 1.7e308)) {  // LIN_I:296 CHECK_GT(segment_time, 0)
      st |= 1u;
      T = 1.0;
    }
    return T;
  };
  auto fixed = [&](int v, int k) { return g.mask[v * HN + k] != 0; };
  auto value = [&](int v, int k, int dim) { return g.values[at<AOS>((size_t)(v * HN + k) * D + dim, rec_v, B, b)]; };
  // H_i[r][c] with r, c in [0, 2h): rows/cols < h belong to the segment's start vertex
  auto Hrc = [&](const double (&pw)[2 * HN - 1], int r, int c) {
    return MTG_H1(r, c) * pw[(r % HN) + (c % HN)];
  };
  double* park = g.scratch + local;  // element s of vertex v at park[((size_t)v * SL + s) * nb]

  double pl[2 * HN - 1], pr[2 * HN - 1];
  double Uprev[HN][HN];  // U_{v-1}: rows free(v-1), cols free(v)
  double G[HN][HN], z[D][HN];
  int f_prev = 0;
  for (int q = 0; q < 2 * HN - 1; ++q) pl[q] = pr[q] = 0.0;

  // ------------------------------------------------------------------ forward
  for (int v = 0; v <= K; ++v) {
    for (int q = 0; q < 2 * HN - 1; ++q) pl[q] = pr[q];
    if (v < K) segment_powers<HN>(seg_time(v), d, pr);
    int fi[HN], f = 0, fn[HN], f_next = 0;
    for (int k = 0; k < HN; ++k)
      if (!fixed(v, k)) fi[f++] = k;
    if (v < K)
      for (int k = 0; k < HN; ++k)
        if (!fixed(v + 1, k)) fn[f_next++] = k;
    double S[HN][HN], rhs[D][HN];
    for (int a = 0; a < f; ++a) {
      const int k = fi[a];
      for (int c = 0; c <= a; ++c) {
        double s = 0.0;
        if (v > 0) s += Hrc(pl, HN + k, HN + fi[c]);
        if (v < K) s += Hrc(pr, k, fi[c]);
        S[a][c] = s;
      }
      for (int dim = 0; dim < D; ++dim) {
        double r = 0.0;
        if (v > 0)
          for (int kk = 0; kk < HN; ++kk) {
            if (fixed(v - 1, kk)) r = fma(-Hrc(pl, HN + k, kk), value(v - 1, kk, dim), r);
            if (fixed(v, kk)) r = fma(-Hrc(pl, HN + k, HN + kk), value(v, kk, dim), r);
          }
        if (v < K)
          for (int kk = 0; kk < HN; ++kk) {
            if (fixed(v, kk)) r = fma(-Hrc(pr, k, kk), value(v, kk, dim), r);
            if (fixed(v + 1, kk)) r = fma(-Hrc(pr, k, HN + kk), value(v + 1, kk, dim), r);
          }
        rhs[dim][a] = r;
      }
    }
    double diag0[HN];  // diagonal of D_v before any elimination: the scale a pivot is judged against
    for (int a = 0; a < f; ++a) diag0[a] = S[a][a];
    // Schur update with the previous vertex: S -= U_{v-1}^T G_{v-1}, rhs -= U_{v-1}^T z_{v-1}
    for (int a = 0; a < f; ++a) {
      for (int c = 0; c <= a; ++c) {
        double s = S[a][c];
        for (int q = 0; q < f_prev; ++q) s = fma(-Uprev[q][a], G[q][c], s);
        S[a][c] = s;
      }
      for (int dim = 0; dim < D; ++dim) {
        double r = rhs[dim][a];
        for (int q = 0; q < f_prev; ++q) r = fma(-Uprev[q][a], z[dim][q], r);
        rhs[dim][a] = r;
      }
    }
    // Cholesky S = L L^T (lower triangle in place), reciprocal diagonal
    double linv[HN];
    for (int j = 0; j < f; ++j) {
      double piv = S[j][j];
      for (int q = 0; q < j; ++q) piv = fma(-S[j][q], S[j][q], piv);
      // a pivot that cancelled to rounding noise is a singular R_pp (cond(R_pp) of a well-posed
      // problem stays below ~1e9)
      if (!(piv > 1e-11 * diag0[j])) {
        st |= 2u;
        piv = 1.0;
      }
      const double rs = rsqrt(piv);
      linv[j] = rs;
      for (int i = j + 1; i < f; ++i) {
        double s = S[i][j];
        for (int q = 0; q < j; ++q) s = fma(-S[i][q], S[j][q], s);
        S[i][j] = s * rs;
      }
    }
    auto solve_in_place = [&](double* x) {  // x <- (L L^T)^-1 x
      for (int i = 0; i < f; ++i) {
        double s = x[i];
        for (int q = 0; q < i; ++q) s = fma(-S[i][q], x[q], s);
        x[i] = s * linv[i];
      }
      for (int i = f - 1; i >= 0; --i) {
        double s = x[i];
        for (int q = i + 1; q < f; ++q) s = fma(-S[q][i], x[q], s);
        x[i] = s * linv[i];
      }
    };
    for (int dim = 0; dim < D; ++dim) {
      solve_in_place(rhs[dim]);
      for (int a = 0; a < f; ++a) z[dim][a] = rhs[dim][a];
    }
    // U_v and G_v = S^-1 U_v
    for (int c = 0; c < f_next; ++c) {
      double col[HN];
      for (int a = 0; a < f; ++a) {
        Uprev[a][c] = Hrc(pr, fi[a], HN + fn[c]);
        col[a] = Uprev[a][c];
      }
      solve_in_place(col);
      for (int a = 0; a < f; ++a) G[a][c] = col[a];
    }
    // park G_v (f x f_next) and z_v (D x f)
    for (int a = 0; a < f; ++a) {
      for (int c = 0; c < f_next; ++c) park[((size_t)v * SL + a * HN + c) * nb] = G[a][c];
      for (int dim = 0; dim < D; ++dim) park[((size_t)v * SL + HN * HN + dim * HN + a) * nb] = z[dim][a];
    }
    f_prev = f;
  }

  // ----------------------------------------------------------------- backward
  // free-constraint offsets per vertex (order of getFreeConstraints: vertex-major, derivative-minor)
  double xe[D][HN], xs[D][HN];  // full derivative vectors (fixed or solved) of vertex v+1 / v
  double x_next[D][HN];         // solved free entries of vertex v+1 (compact)
  int f_next = 0, off = g.n_free;
  double cost_acc = 0.0;
  const size_t rec_free = (size_t)D * g.n_free;
  for (int v = K; v >= 0; --v) {
    int fi[HN], f = 0;
    for (int k = 0; k < HN; ++k)
      if (!fixed(v, k)) fi[f++] = k;
    off -= f;
    double x[D][HN];
    for (int dim = 0; dim < D; ++dim)
      for (int a = 0; a < f; ++a) {
        double s = park[((size_t)v * SL + HN * HN + dim * HN + a) * nb];
        for (int c = 0; c < f_next; ++c) s = fma(-park[((size_t)v * SL + a * HN + c) * nb], x_next[dim][c], s);
        x[dim][a] = s;
      }
    for (int dim = 0; dim < D; ++dim) {
      int a = 0;
      for (int k = 0; k < HN; ++k) {
        if (fixed(v, k)) {
          xs[dim][k] = value(v, k, dim);
        } else {
          xs[dim][k] = x[dim][a];
          if (g.free_out) g.free_out[at<AOS>((size_t)dim * g.n_free + off + a, rec_free, B, b)] = x[dim][a];
          ++a;
        }
      }
    }
    if (v < K) cost_acc += emit_segment<HN, D, AOS>(p, v, b, true, seg_time(v), xs, xe);
    for (int dim = 0; dim < D; ++dim) {
      for (int k = 0; k < HN; ++k) xe[dim][k] = xs[dim][k];
      for (int a = 0; a < f; ++a) x_next[dim][a] = x[dim][a];
    }
    f_next = f;
  }
  if (p.cost) p.cost[b] = 0.5 * cost_acc;
  if (p.status) p.status[b] = st;
}

}  // namespace mtg
#endif

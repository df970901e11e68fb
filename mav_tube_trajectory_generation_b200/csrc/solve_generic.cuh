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
// parked in a global scratch, slot-major / batch-minor (coalesced). The kernel below works on the
// EMBEDDED system (every vertex a full h x h block, fixed entries as identity rows), see its comment.
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

// Register-resident form: every vertex is treated as a FULL h x h block. A fixed derivative (v, k) keeps its
// place in the block as an identity row / column (its known value moves to the right-hand sides of the rows it
// couples to), so block sizes and every array index are compile-time constants, the mask only supplies uniform
// predicates (it is shared by the batch: no divergence), and S, U, G, z and the right-hand sides live in
// registers instead of local memory. The factorisation of the free part is unchanged: the identity rows
// decouple exactly.
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
  // bit k of fx(v): derivative k of vertex v is fixed (out-of-range vertices: nothing fixed, never used)
  auto fx = [&](int v) -> unsigned {
    unsigned m = 0;
    if (v >= 0 && v <= K) {
#pragma unroll
      for (int k = 0; k < HN; ++k) m |= (g.mask[v * HN + k] != 0 ? 1u : 0u) << k;
    }
    return m;
  };
  auto load_values = [&](int v, unsigned m, double (&val)[D][HN]) {
#pragma unroll
    for (int dim = 0; dim < D; ++dim)
#pragma unroll
      for (int k = 0; k < HN; ++k)
        val[dim][k] = ((m >> k) & 1u) ? g.values[at<AOS>((size_t)(v * HN + k) * D + dim, rec_v, B, b)] : 0.0;
  };
  double* park = g.scratch + local;  // element s of vertex v at park[((size_t)v * SL + s) * nb]

  double pl[2 * HN - 1], pr[2 * HN - 1];
  double U[HN][HN];   // masked U_{v-1}: rows = derivatives of v-1, cols = derivatives of v
  double G[HN][HN], z[D][HN];
  double val_prev[D][HN], val_cur[D][HN], val_next[D][HN];
#pragma unroll
  for (int q = 0; q < 2 * HN - 1; ++q) pl[q] = pr[q] = 0.0;
#pragma unroll
  for (int i = 0; i < HN; ++i) {
#pragma unroll
    for (int c = 0; c < HN; ++c) U[i][c] = G[i][c] = 0.0;
#pragma unroll
    for (int dim = 0; dim < D; ++dim) z[dim][i] = val_prev[dim][i] = 0.0;
  }
  unsigned m_prev = 0, m_cur = fx(0), m_next = fx(1);
  load_values(0, m_cur, val_cur);
  load_values(1, m_next, val_next);

  // ------------------------------------------------------------------ forward
  for (int v = 0; v <= K; ++v) {
#pragma unroll
    for (int q = 0; q < 2 * HN - 1; ++q) pl[q] = pr[q];
    if (v < K) segment_powers<HN>(seg_time(v), d, pr);
    const bool has_l = v > 0, has_r = v < K;
    // H_{v-1}[r][c] = H1[r][c] pl[r%h + c%h], H_v likewise with pr
    double S[HN][HN], rhs[D][HN];
#pragma unroll
    for (int k = 0; k < HN; ++k) {
#pragma unroll
      for (int c = 0; c <= k; ++c) {
        double s = 0.0;
        if (has_l) s += MTG_H1(HN + k, HN + c) * pl[k + c];
        if (has_r) s += MTG_H1(k, c) * pr[k + c];
        S[k][c] = s;
      }
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
        double r = 0.0;
#pragma unroll
        for (int kk = 0; kk < HN; ++kk) {
          // known values of the three vertices this row couples to (val is 0 where nothing is fixed)
          if (has_l) {
            r = fma(-MTG_H1(HN + k, kk) * pl[k + kk], val_prev[dim][kk], r);
            r = fma(-MTG_H1(HN + k, HN + kk) * pl[k + kk], val_cur[dim][kk], r);
          }
          if (has_r) {
            r = fma(-MTG_H1(k, kk) * pr[k + kk], val_cur[dim][kk], r);
            r = fma(-MTG_H1(k, HN + kk) * pr[k + kk], val_next[dim][kk], r);
          }
        }
        rhs[dim][k] = r;
      }
    }
    // fixed derivatives of this vertex: identity row / column, right-hand side = the value
#pragma unroll
    for (int k = 0; k < HN; ++k) {
      const bool fk = (m_cur >> k) & 1u;
#pragma unroll
      for (int c = 0; c <= k; ++c) {
        const bool fc = (m_cur >> c) & 1u;
        if (fk || fc) S[k][c] = (k == c) ? 1.0 : 0.0;
      }
      if (fk) {
#pragma unroll
        for (int dim = 0; dim < D; ++dim) rhs[dim][k] = val_cur[dim][k];
      }
    }
    double diag0[HN];  // diagonal before any elimination: the scale a pivot is judged against
#pragma unroll
    for (int a = 0; a < HN; ++a) diag0[a] = S[a][a];
    // Schur update with the previous vertex (U is masked: fixed rows / columns are zero)
    if (has_l) {
#pragma unroll
      for (int a = 0; a < HN; ++a) {
#pragma unroll
        for (int c = 0; c <= a; ++c) {
          double s = S[a][c];
#pragma unroll
          for (int q = 0; q < HN; ++q) s = fma(-U[q][a], G[q][c], s);
          S[a][c] = s;
        }
#pragma unroll
        for (int dim = 0; dim < D; ++dim) {
          double r = rhs[dim][a];
#pragma unroll
          for (int q = 0; q < HN; ++q) r = fma(-U[q][a], z[dim][q], r);
          rhs[dim][a] = r;
        }
      }
    }
    // Cholesky S = L L^T (lower triangle in place), reciprocal diagonal
    double linv[HN];
#pragma unroll
    for (int j = 0; j < HN; ++j) {
      double piv = S[j][j];
#pragma unroll
      for (int q = 0; q < j; ++q) piv = fma(-S[j][q], S[j][q], piv);
      // a pivot that cancelled to rounding noise is a singular R_pp (cond(R_pp) of a well-posed
      // problem stays below ~1e9)
      if (!(piv > 1e-11 * diag0[j])) {
        st |= 2u;
        piv = 1.0;
      }
      const double rs = rsqrt(piv);
      linv[j] = rs;
#pragma unroll
      for (int i = j + 1; i < HN; ++i) {
        double s = S[i][j];
#pragma unroll
        for (int q = 0; q < j; ++q) s = fma(-S[i][q], S[j][q], s);
        S[i][j] = s * rs;
      }
    }
    auto solve_in_place = [&](double (&x)[HN]) {  // x <- (L L^T)^-1 x
#pragma unroll
      for (int i = 0; i < HN; ++i) {
        double s = x[i];
#pragma unroll
        for (int q = 0; q < i; ++q) s = fma(-S[i][q], x[q], s);
        x[i] = s * linv[i];
      }
#pragma unroll
      for (int i = HN - 1; i >= 0; --i) {
        double s = x[i];
#pragma unroll
        for (int q = i + 1; q < HN; ++q) s = fma(-S[q][i], x[q], s);
        x[i] = s * linv[i];
      }
    };
#pragma unroll
    for (int dim = 0; dim < D; ++dim) {
      solve_in_place(rhs[dim]);
#pragma unroll
      for (int a = 0; a < HN; ++a) z[dim][a] = rhs[dim][a];
    }
    // masked U_v = H_v[start, end] and G_v = S^-1 U_v
#pragma unroll
    for (int c = 0; c < HN; ++c) {
      double col[HN];
#pragma unroll
      for (int a = 0; a < HN; ++a) {
        const bool live = has_r && !((m_cur >> a) & 1u) && !((m_next >> c) & 1u);
        U[a][c] = live ? MTG_H1(a, HN + c) * pr[a + c] : 0.0;
        col[a] = U[a][c];
      }
      solve_in_place(col);
#pragma unroll
      for (int a = 0; a < HN; ++a) G[a][c] = col[a];
    }
    // park G_v and z_v
#pragma unroll
    for (int a = 0; a < HN; ++a) {
#pragma unroll
      for (int c = 0; c < HN; ++c) park[((size_t)v * SL + a * HN + c) * nb] = G[a][c];
#pragma unroll
      for (int dim = 0; dim < D; ++dim) park[((size_t)v * SL + HN * HN + dim * HN + a) * nb] = z[dim][a];
    }
    // slide the window of known values
    m_prev = m_cur;
    m_cur = m_next;
    m_next = fx(v + 2);
#pragma unroll
    for (int dim = 0; dim < D; ++dim)
#pragma unroll
      for (int k = 0; k < HN; ++k) {
        val_prev[dim][k] = val_cur[dim][k];
        val_cur[dim][k] = val_next[dim][k];
      }
    if (v + 2 <= K) {
      load_values(v + 2, m_next, val_next);
    } else {
#pragma unroll
      for (int dim = 0; dim < D; ++dim)
#pragma unroll
        for (int k = 0; k < HN; ++k) val_next[dim][k] = 0.0;
    }
  }
  (void)m_prev;

  // ----------------------------------------------------------------- backward: x_v = z_v - G_v x_{v+1}
  double xe[D][HN], xs[D][HN];  // full derivative vectors of vertex v+1 / v
#pragma unroll
  for (int dim = 0; dim < D; ++dim)
#pragma unroll
    for (int k = 0; k < HN; ++k) xe[dim][k] = 0.0;
  int off = g.n_free;
  double cost_acc = 0.0;
  const size_t rec_free = (size_t)D * g.n_free;
  for (int v = K; v >= 0; --v) {
    const unsigned m = fx(v);
    const int f = HN - __popc(m);
    off -= f;
#pragma unroll
    for (int dim = 0; dim < D; ++dim)
#pragma unroll
      for (int a = 0; a < HN; ++a) {
        double s = park[((size_t)v * SL + HN * HN + dim * HN + a) * nb];
        if (v < K) {
#pragma unroll
          for (int c = 0; c < HN; ++c) s = fma(-park[((size_t)v * SL + a * HN + c) * nb], xe[dim][c], s);
        }
        xs[dim][a] = s;
      }
    if (g.free_out) {
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
        int a = 0;
#pragma unroll
        for (int k = 0; k < HN; ++k)
          if (!((m >> k) & 1u)) {
            g.free_out[at<AOS>((size_t)dim * g.n_free + off + a, rec_free, B, b)] = xs[dim][k];
            ++a;
          }
      }
    }
    if (v < K) cost_acc += emit_segment<HN, D, AOS>(p, v, b, true, seg_time(v), xs, xe);
#pragma unroll
    for (int dim = 0; dim < D; ++dim)
#pragma unroll
      for (int k = 0; k < HN; ++k) xe[dim][k] = xs[dim][k];
  }
  if (p.cost) p.cost[b] = 0.5 * cost_acc;
  if (p.status) p.status[b] = st;
}

}  // namespace mtg
#endif

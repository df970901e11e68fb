// solve_canonical — batched closed-form min-derivative solve, one THREAD per
// trajectory, for the constraint pattern of createRandomVertices
// (reference src/vertex.cpp:27-82): first/last vertex fix derivatives 0..h-1,
// interior vertices fix position only.
//
// Replaces per trajectory (reference include/.../impl/polynomial_optimization_linear_impl.h):
//   updateSegmentTimes :277-304 (Q :557-573, A :101-111, A^-1 :132-169)
//   setupConstraintReorderingMatrix :171-252     constructR :306-335
//   solveLinear :337-379     updateSegmentsFromCompactConstraints :254-275
//   computeCost :113-130
//
// Math: with H_i = T_i^(1-2d) S_i H1 S_i the normal matrix R_pp of the
// free derivatives (k = 1..h-1 at vertices 1..K-1) is block tridiagonal with
// (h-1)x(h-1) blocks
//   D_v = H_{v-1}[end,end] + H_v[start,start],   U_v = H_v[start,end]
// and is solved by a block Thomas sweep (Cholesky of each Schur block):
//   S_v = D_v - U_{v-1}^T G_{v-1},  G_v = S_v^-1 U_v,  z_v = S_v^-1 (b_v - U_{v-1}^T z_{v-1})
//   x_v = z_v - G_v x_{v+1}
// The sweep state (G_v, z_v for v = 1..K-2) lives in shared memory,
// slot-major / thread-minor (bank-conflict free); everything else in registers.
// Coefficients come from the scaled constant inverse:
//   c_j = d_j / j! (j < h),   c_j = T^-j sum_m Ainv1[j][m] (T^alpha_m d_m)  (j >= h)
// and the cost from the same scaled endpoint vector: 0.5 T^(1-2d) dhat^T H1 dhat.
#ifndef MTG_SOLVE_CANONICAL_CUH_
#define MTG_SOLVE_CANONICAL_CUH_

#include <stdint.h>

#include "device_tables.cuh"

namespace mtg {

struct SolveCanonicalParams {
  const double* __restrict__ positions;        // [K+1][D][B]
  const double* __restrict__ end_derivatives;  // [2][h-1][D][B] or nullptr
  const double* __restrict__ seg_times;        // [K][B]
  double* __restrict__ coeffs;                 // [K][D][N][B]
  double* __restrict__ cost;                   // [B] or nullptr
  double* __restrict__ free_constraints;       // [K-1][h-1][D][B] or nullptr
  uint32_t* __restrict__ status;               // [B] or nullptr
  int B;   // leading dimension of every tensor
  int b0;  // first trajectory handled by this launch
  int nb;  // number of trajectories handled by this launch
  int K;
  int derivative;
};

template <int HN>
__device__ __forceinline__ void segment_powers(double T, int derivative, double (&pw)[2 * HN - 1]) {
  // pw[q] = T^(q + 1 - 2d), q = alpha_r + alpha_c in [0, 2h-2]
  const int e0 = 1 - 2 * derivative;
  double base = 1.0;
  if (e0 >= 0) {
    for (int i = 0; i < e0; ++i) base *= T;
  } else {
    const double u = 1.0 / T;
    for (int i = 0; i < -e0; ++i) base *= u;
  }
  pw[0] = base;
#pragma unroll
  for (int q = 1; q < 2 * HN - 1; ++q) pw[q] = pw[q - 1] * T;
}

#define MTG_H1(r, c) c_tab.H1[(r) * MTG_TAB_LD + (c)]
#define MTG_AI(r, c) c_tab.Ainv1[(r) * MTG_TAB_LD + (c)]

// Writes the N coefficients of every dimension of one segment and returns the
// segment's cost contribution  T^(1-2d) * sum_dim dhat^T H1 dhat  (without 1/2).
// ds / de: derivative values 0..h-1 at the segment start / end, per dimension.
template <int HN, int D>
__device__ __forceinline__ double emit_segment(const SolveCanonicalParams& p, int seg, int b, double T,
                                               const double (&ds)[D][HN], const double (&de)[D][HN]) {
  constexpr int N = 2 * HN;
  double tp[HN];  // T^m
  tp[0] = 1.0;
#pragma unroll
  for (int m = 1; m < HN; ++m) tp[m] = tp[m - 1] * T;
  const double u = 1.0 / T;
  double uh = u;  // u^HN
#pragma unroll
  for (int m = 1; m < HN; ++m) uh *= u;
  double quad = 0.0;
#pragma unroll
  for (int dim = 0; dim < D; ++dim) {
    double dh[N];
#pragma unroll
    for (int m = 0; m < HN; ++m) {
      dh[m] = tp[m] * ds[dim][m];
      dh[HN + m] = tp[m] * de[dim][m];
    }
    double* out = p.coeffs + ((size_t)(seg * D + dim) * N) * p.B + b;
#pragma unroll
    for (int j = 0; j < HN; ++j) out[(size_t)j * p.B] = ds[dim][j] * c_tab.inv_factorial[j];
    double us = uh;
#pragma unroll
    for (int j = HN; j < N; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int m = 0; m < N; ++m) acc = fma(MTG_AI(j, m), dh[m], acc);
      out[(size_t)j * p.B] = acc * us;
      us *= u;
    }
    // dhat^T H1 dhat using symmetry
#pragma unroll
    for (int r = 0; r < N; ++r) {
      double row = 0.5 * MTG_H1(r, r) * dh[r];
#pragma unroll
      for (int c = r + 1; c < N; ++c) row = fma(MTG_H1(r, c), dh[c], row);
      quad = fma(2.0 * dh[r], row, quad);
    }
  }
  // T^(1-2d)
  double s = 1.0;
  const int e0 = 1 - 2 * p.derivative;
  if (e0 >= 0) {
    for (int i = 0; i < e0; ++i) s *= T;
  } else {
    for (int i = 0; i < -e0; ++i) s *= u;
  }
  return quad * s;
}

template <int HN, int D>
__global__ void __launch_bounds__(128) solve_canonical_kernel(const SolveCanonicalParams p) {
  constexpr int NF = HN - 1;             // free derivatives per interior vertex
  constexpr int SLOTS = NF * NF + NF * D;  // G_v and z_v
  extern __shared__ double smem[];
  const int tid = threadIdx.x;
  const int nt = blockDim.x;
  const int local = blockIdx.x * nt + tid;
  if (local >= p.nb) return;
  const int b = p.b0 + local;
  const int K = p.K;
  const size_t B = (size_t)p.B;
  const int d = p.derivative;
  uint32_t st = 0;

  auto pos = [&](int v, int dim) { return p.positions[((size_t)v * D + dim) * B + b]; };
  auto seg_time = [&](int i) {
    double T = p.seg_times[(size_t)i * B + b];
    if (!(T > 0.0) || !(T < 1.7e308)) {  // LIN_I:296 CHECK_GT(segment_time, 0)
      st |= 1u;
      T = 1.0;
    }
    return T;
  };

  // endpoint derivative constraints (zero = makeStartOrEnd)
  double sd[D][HN], ed[D][HN];
#pragma unroll
  for (int dim = 0; dim < D; ++dim) {
    sd[dim][0] = pos(0, dim);
    ed[dim][0] = pos(K, dim);
#pragma unroll
    for (int m = 1; m < HN; ++m) {
      sd[dim][m] = 0.0;
      ed[dim][m] = 0.0;
    }
  }
  const bool have_end = (p.end_derivatives != nullptr);
  if (have_end) {
#pragma unroll
    for (int dim = 0; dim < D; ++dim)
#pragma unroll
      for (int m = 1; m < HN; ++m) {
        sd[dim][m] = p.end_derivatives[((size_t)(0 * NF + (m - 1)) * D + dim) * B + b];
        ed[dim][m] = p.end_derivatives[((size_t)(1 * NF + (m - 1)) * D + dim) * B + b];
      }
  }

  double cost_acc = 0.0;

  if (K == 1) {
    const double T = seg_time(0);
    cost_acc = emit_segment<HN, D>(p, 0, b, T, sd, ed);
  } else {
    // ------------------------------------------------------------ forward
    double pl[2 * HN - 1], pr[2 * HN - 1];
    double U[NF][NF];   // U_{v-1} on entry of step v (rows: vertex v-1, cols: vertex v)
    double G[NF][NF];   // G_{v-1}
    double z[D][NF];    // z_{v-1}
    double p_prev[D], p_cur[D], p_next[D];
    {
      const double T0 = seg_time(0);
      segment_powers<HN>(T0, d, pr);
    }
#pragma unroll
    for (int dim = 0; dim < D; ++dim) {
      p_cur[dim] = sd[dim][0];
      p_next[dim] = pos(1, dim);
    }
#pragma unroll 1
    for (int v = 1; v <= K - 1; ++v) {
#pragma unroll
      for (int q = 0; q < 2 * HN - 1; ++q) pl[q] = pr[q];
      segment_powers<HN>(seg_time(v), d, pr);
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
        p_prev[dim] = p_cur[dim];
        p_cur[dim] = p_next[dim];
        p_next[dim] = pos(v + 1, dim);
      }
      // Schur block (lower triangle) and right-hand sides
      double S[NF][NF];
      double r[D][NF];
#pragma unroll
      for (int k = 0; k < NF; ++k) {
#pragma unroll
        for (int kk = 0; kk <= k; ++kk) {
          double s = MTG_H1(HN + 1 + k, HN + 1 + kk) * pl[k + kk + 2];
          s = fma(MTG_H1(1 + k, 1 + kk), pr[k + kk + 2], s);
          S[k][kk] = s;
        }
        const double a_self = fma(MTG_H1(HN + 1 + k, HN), pl[k + 1], MTG_H1(1 + k, 0) * pr[k + 1]);
        const double a_prev = MTG_H1(HN + 1 + k, 0) * pl[k + 1];
        const double a_next = MTG_H1(1 + k, HN) * pr[k + 1];
#pragma unroll
        for (int dim = 0; dim < D; ++dim)
          r[dim][k] = -fma(a_self, p_cur[dim], fma(a_prev, p_prev[dim], a_next * p_next[dim]));
      }
      if (have_end) {
        if (v == 1) {
#pragma unroll
          for (int k = 0; k < NF; ++k)
#pragma unroll
            for (int m = 1; m < HN; ++m) {
              const double a = MTG_H1(HN + 1 + k, m) * pl[k + 1 + m];
#pragma unroll
              for (int dim = 0; dim < D; ++dim) r[dim][k] = fma(-a, sd[dim][m], r[dim][k]);
            }
        }
        if (v == K - 1) {
#pragma unroll
          for (int k = 0; k < NF; ++k)
#pragma unroll
            for (int m = 1; m < HN; ++m) {
              const double a = MTG_H1(1 + k, HN + m) * pr[k + 1 + m];
#pragma unroll
              for (int dim = 0; dim < D; ++dim) r[dim][k] = fma(-a, ed[dim][m], r[dim][k]);
            }
        }
      }
      if (v > 1) {
#pragma unroll
        for (int k = 0; k < NF; ++k) {
#pragma unroll
          for (int kk = 0; kk <= k; ++kk) {
            double s = S[k][kk];
#pragma unroll
            for (int j = 0; j < NF; ++j) s = fma(-U[j][k], G[j][kk], s);
            S[k][kk] = s;
          }
#pragma unroll
          for (int dim = 0; dim < D; ++dim) {
            double s = r[dim][k];
#pragma unroll
            for (int j = 0; j < NF; ++j) s = fma(-U[j][k], z[dim][j], s);
            r[dim][k] = s;
          }
        }
      }
      // Cholesky S = L L^T, diagonal kept as its reciprocal
      double linv[NF];
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        double piv = S[j][j];
#pragma unroll
        for (int q = 0; q < j; ++q) piv = fma(-S[j][q], S[j][q], piv);
        if (!(piv > 0.0)) {
          st |= 2u;
          piv = 1.0;
        }
        const double rs2 = rsqrt(piv);
        linv[j] = rs2;
#pragma unroll
        for (int i = j + 1; i < NF; ++i) {
          double s = S[i][j];
#pragma unroll
          for (int q = 0; q < j; ++q) s = fma(-S[i][q], S[j][q], s);
          S[i][j] = s * rs2;
        }
      }
      // z_v = S^-1 r   (L y = r ; L^T z = y)
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
#pragma unroll
        for (int i = 0; i < NF; ++i) {
          double s = r[dim][i];
#pragma unroll
          for (int q = 0; q < i; ++q) s = fma(-S[i][q], z[dim][q], s);
          z[dim][i] = s * linv[i];
        }
#pragma unroll
        for (int i = NF - 1; i >= 0; --i) {
          double s = z[dim][i];
#pragma unroll
          for (int q = i + 1; q < NF; ++q) s = fma(-S[q][i], z[dim][q], s);
          z[dim][i] = s * linv[i];
        }
      }
      if (v < K - 1) {
        // U_v and G_v = S^-1 U_v
#pragma unroll
        for (int k = 0; k < NF; ++k)
#pragma unroll
          for (int kk = 0; kk < NF; ++kk) U[k][kk] = MTG_H1(1 + k, HN + 1 + kk) * pr[k + kk + 2];
#pragma unroll
        for (int c = 0; c < NF; ++c) {
#pragma unroll
          for (int i = 0; i < NF; ++i) {
            double s = U[i][c];
#pragma unroll
            for (int q = 0; q < i; ++q) s = fma(-S[i][q], G[q][c], s);
            G[i][c] = s * linv[i];
          }
#pragma unroll
          for (int i = NF - 1; i >= 0; --i) {
            double s = G[i][c];
#pragma unroll
            for (int q = i + 1; q < NF; ++q) s = fma(-S[q][i], G[q][c], s);
            G[i][c] = s * linv[i];
          }
        }
        // park G_v, z_v
        double* slot = smem + (size_t)(v - 1) * SLOTS * nt + tid;
#pragma unroll
        for (int i = 0; i < NF; ++i)
#pragma unroll
          for (int c = 0; c < NF; ++c) slot[(size_t)(i * NF + c) * nt] = G[i][c];
#pragma unroll
        for (int dim = 0; dim < D; ++dim)
#pragma unroll
          for (int i = 0; i < NF; ++i) slot[(size_t)(NF * NF + dim * NF + i) * nt] = z[dim][i];
      }
    }
    // ----------------------------------------------------------- backward
    // on exit: z = x_{K-1}; p_cur = pos(K-1), p_next = pos(K)
    double xs[D][HN], xe[D][HN];
#pragma unroll
    for (int dim = 0; dim < D; ++dim) {
      xs[dim][0] = p_cur[dim];
#pragma unroll
      for (int i = 0; i < NF; ++i) xs[dim][1 + i] = z[dim][i];
    }
    if (p.free_constraints) {
#pragma unroll
      for (int dim = 0; dim < D; ++dim)
#pragma unroll
        for (int i = 0; i < NF; ++i)
          p.free_constraints[((size_t)((K - 2) * NF + i) * D + dim) * B + b] = z[dim][i];
    }
    cost_acc += emit_segment<HN, D>(p, K - 1, b, seg_time(K - 1), xs, ed);
#pragma unroll 1
    for (int v = K - 2; v >= 1; --v) {
#pragma unroll
      for (int dim = 0; dim < D; ++dim)
#pragma unroll
        for (int m = 0; m < HN; ++m) xe[dim][m] = xs[dim][m];
      const double* slot = smem + (size_t)(v - 1) * SLOTS * nt + tid;
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
        xs[dim][0] = pos(v, dim);
#pragma unroll
        for (int i = 0; i < NF; ++i) {
          double s = slot[(size_t)(NF * NF + dim * NF + i) * nt];
#pragma unroll
          for (int c = 0; c < NF; ++c) s = fma(-slot[(size_t)(i * NF + c) * nt], xe[dim][1 + c], s);
          xs[dim][1 + i] = s;
        }
      }
      if (p.free_constraints) {
#pragma unroll
        for (int dim = 0; dim < D; ++dim)
#pragma unroll
          for (int i = 0; i < NF; ++i)
            p.free_constraints[((size_t)((v - 1) * NF + i) * D + dim) * B + b] = xs[dim][1 + i];
      }
      cost_acc += emit_segment<HN, D>(p, v, b, seg_time(v), xs, xe);
    }
    cost_acc += emit_segment<HN, D>(p, 0, b, seg_time(0), sd, xs);
  }
  if (p.cost) p.cost[b] = 0.5 * cost_acc;
  if (p.status) p.status[b] = st;
}

}  // namespace mtg
#endif

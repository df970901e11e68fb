// solve_canonical — batched closed-form min-derivative solve for the constraint
// pattern of createRandomVertices (reference src/vertex.cpp:27-82): first/last
// vertex fix derivatives 0..h-1, interior vertices fix position only.
//
// Replaces per trajectory (reference include/.../impl/polynomial_optimization_linear_impl.h):
//   updateSegmentTimes :277-304 (Q :557-573, A :101-111, A^-1 :132-169)
//   setupConstraintReorderingMatrix :171-252     constructR :306-335
//   solveLinear :337-379     updateSegmentsFromCompactConstraints :254-275
//   computeCost :113-130
//
// Math. With H_i = T_i^(1-2d) S_i H1 S_i the normal matrix R_pp of
// the free derivatives (orders 1..h-1 at vertices 1..K-1) is block tridiagonal,
//   D_v = H_{v-1}[end,end] + H_v[start,start],   U_v = H_v[start,end],
// and the right-hand side only needs position DIFFERENCES (H annihilates constant
// offsets: H[r][0] = -H[r][h]), so the solve is translation invariant by construction.
//
// Mapping: TWO LANES PER TRAJECTORY ("twisted" block factorisation). Lane A
// eliminates vertices 1..m-1 top-down, lane B eliminates K-1..m+1 bottom-up, both
// with the same block-Thomas recurrences
//   S_j = D_j - U_{j-1}^T G_{j-1},  G_j = S_j^-1 U_j,  z_j = S_j^-1 (r_j - U_{j-1}^T z_{j-1}).
// Lane B runs the SAME code on the time-reversed trajectory: reversing time swaps
// segment start/end and negates odd derivatives, H1[pi r][pi c] = (-1)^(a_r+a_c) H1[r][c],
// so the mirrored chain has exactly the original form (no divergence between the
// lanes, constant tables read uniformly). The lanes exchange their Schur
// contributions to the meeting vertex m with one round of shuffles, both solve the
// (h-1)x(h-1) meeting system, then each back-substitutes x_j = z_j - G_j x_{j+1} over
// its half and writes the coefficients and cost of its half of the segments.
// The parked (G_j, z_j) live in shared memory, slot-major / thread-minor
// (conflict free); half the state per lane doubles the resident warps.
//
// Coefficients:  c_j = d_j / j! (j < h),
//                c_j = T^-j [Ainv1[j][h] (p_e - p_s) + sum_{m>=1} Ainv1[j][m] T^m d_m + ...]
// Cost: 0.5 T^(1-2d) |W dhat|^2 with H1 = W^T W (tables.cpp): a sum of squares; a
// direct dhat^T H1 dhat cancels ~6 digits.
#ifndef MTG_SOLVE_CANONICAL_CUH_
#define MTG_SOLVE_CANONICAL_CUH_

#include <stdint.h>

#include "argmin.cuh"
#include "device_tables.cuh"

namespace mtg {

struct SolveCanonicalParams {
  const double* __restrict__ positions;        // elem (v*D + dim),                 rec (K+1)*D
  const double* __restrict__ end_derivatives;  // elem ((side*(h-1) + m-1)*D + dim), rec 2*(h-1)*D; or nullptr
  const double* __restrict__ seg_times;        // elem i,                            rec K
  double* __restrict__ coeffs;                 // elem ((i*D + dim)*N + j),          rec K*D*N
  double* __restrict__ cost;                   // [B] or nullptr
  double* __restrict__ free_constraints;       // elem ((dim*(K-1) + v-1)*(h-1) + k-1), rec D*(K-1)*(h-1); or nullptr
  uint32_t* __restrict__ status;               // [B] or nullptr
  int B;   // leading dimension (SoA) / number of records
  int b0;  // first trajectory handled by this launch
  int nb;  // number of trajectories handled by this launch
  int K;
  int derivative;
  int vec_ok;  // AoS only: coeffs is 16-byte aligned -> double2 stores
  // fused candidate argmin (mtg_solve_argmin_batch; best_out == nullptr: off): every CTA folds the best
  // {cost, global index} of its trajectories into the running pair *best_out
  int overlap = 0;  // 1: launched with programmatic stream serialization (mtg_set_solve_overlap): the grid may start
                    // while the previous kernel of the stream drains and waits for it before its first global store
  unsigned* best_lock = nullptr;        // zero when free
  void* best_out = nullptr;             // Best*: the running best, {+inf, -1} = nothing yet
  long long best_offset = 0;            // global index of trajectory 0 of the batch
};

// element `elem` of record `b`: SoA = batch innermost, AoS = record-contiguous
template <bool AOS>
__device__ __forceinline__ size_t at(size_t elem, size_t rec, size_t B, size_t b) {
  return AOS ? b * rec + elem : elem * B + b;
}

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
#define MTG_W(r, c) c_tab.W[(r) * MTG_TAB_LD + (c)]
#define MTG_LT(r, c) c_tab.Lt[(r) * MTG_TAB_LD + (c)]

// Writes the N coefficients of every dimension of one segment (original
// orientation: ds at the segment start, de at its end) and returns the segment's
// cost contribution  T^(1-2d) * sum_dim |W dhat|^2  (without the 1/2).
// DT >= 0: the cost derivative as a compile-time constant (the triangular cost loop then has no
// predicated-off slots); DT = -1: taken from p.derivative.
template <int HN, int D, bool AOS, int DT = -1>
__device__ __forceinline__ double emit_segment(const SolveCanonicalParams& p, int seg, int b, bool active,
                                               double T, const double (&ds)[D][HN],
                                               const double (&de)[D][HN]) {
  constexpr int N = 2 * HN;
  double tp[HN];  // T^m
  tp[0] = 1.0;
#pragma unroll
  for (int m = 1; m < HN; ++m) tp[m] = tp[m - 1] * T;
  const double u = 1.0 / T;
  double uh = u;  // u^HN
#pragma unroll
  for (int m = 1; m < HN; ++m) uh *= u;
  const int dcost = DT >= 0 ? DT : p.derivative;
  const int nq = N - dcost;
  double quad = 0.0;
  const size_t rec = (size_t)p.K * D * N;
#pragma unroll
  for (int dim = 0; dim < D; ++dim) {
    double dh[N];  // scaled endpoint derivatives; slots 0 and HN unused (position enters as a difference)
    const double dlt = de[dim][0] - ds[dim][0];
#pragma unroll
    for (int m = 1; m < HN; ++m) {
      dh[m] = tp[m] * ds[dim][m];
      dh[HN + m] = tp[m] * de[dim][m];
    }
    // chat_j = c_j T^j: the scaled coefficients. Lower half straight from the start derivatives,
    // upper half through A(1)^-1; c_j = chat_j T^-j.
    double c[N], chat[N];
#pragma unroll
    for (int j = 0; j < HN; ++j) {
      c[j] = ds[dim][j] * c_tab.inv_factorial[j];
      chat[j] = (j == 0 ? ds[dim][0] : dh[j]) * c_tab.inv_factorial[j];
    }
    double us = uh;
#pragma unroll
    for (int j = HN; j < N; ++j) {
      double acc = MTG_AI(j, HN) * dlt;
#pragma unroll
      for (int m = 1; m < HN; ++m) {
        acc = fma(MTG_AI(j, m), dh[m], acc);
        acc = fma(MTG_AI(j, HN + m), dh[HN + m], acc);
      }
      chat[j] = acc;
      c[j] = acc * us;
      us *= u;
    }
    if (active && p.coeffs) {  // coeffs == nullptr: cost-only solve (candidate sweeps)
      if (AOS) {
        double* out = p.coeffs + (size_t)b * rec + (size_t)(seg * D + dim) * N;
        if (p.vec_ok) {
#pragma unroll
          for (int j = 0; j < N; j += 2) *reinterpret_cast<double2*>(out + j) = make_double2(c[j], c[j + 1]);
        } else {
#pragma unroll
          for (int j = 0; j < N; ++j) out[j] = c[j];
        }
      } else {
        double* out = p.coeffs + ((size_t)(seg * D + dim) * N) * p.B + b;
#pragma unroll
        for (int j = 0; j < N; ++j) out[(size_t)j * p.B] = c[j];
      }
    }
    // cost: |W dhat|^2 = |Lt chat[d..]|^2 — a sum of squares (no cancellation across terms), and the
    // triangular Lt costs (N-d)(N-d+1)/2 multiply-adds instead of the (N-d)(N-1) of W
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (i < nq) {
        double w = 0.0;
#pragma unroll
        for (int j = i; j < N; ++j)  // chat index j = dcost + a with a >= i
          if (j >= dcost + i) w = fma(c_tab.Lt[i * MTG_TAB_LD + (j - dcost)], chat[j], w);
        quad = fma(w, w, quad);
      }
    }
  }
  double s = 1.0;  // T^(1-2d)
  const int e0 = 1 - 2 * dcost;
  if (e0 >= 0) {
    for (int i = 0; i < e0; ++i) s *= T;
  } else {
    for (int i = 0; i < -e0; ++i) s *= u;
  }
  return quad * s;
}

// In-place Cholesky of the lower triangle of S (reciprocal diagonal in linv).
template <int NF>
__device__ __forceinline__ void chol_lower(double (&S)[NF][NF], double (&linv)[NF], uint32_t& st) {
#pragma unroll
  for (int j = 0; j < NF; ++j) {
    double piv = S[j][j];
#pragma unroll
    for (int q = 0; q < j; ++q) piv = fma(-S[j][q], S[j][q], piv);
    if (!(piv > 0.0)) {
      st |= 2u;
      piv = 1.0;
    }
    const double rs = rsqrt(piv);
    linv[j] = rs;
#pragma unroll
    for (int i = j + 1; i < NF; ++i) {
      double s = S[i][j];
#pragma unroll
      for (int q = 0; q < j; ++q) s = fma(-S[i][q], S[j][q], s);
      S[i][j] = s * rs;
    }
  }
}

// x <- (L L^T)^-1 x
template <int NF>
__device__ __forceinline__ void chol_solve(const double (&L)[NF][NF], const double (&linv)[NF], double (&x)[NF]) {
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    double s = x[i];
#pragma unroll
    for (int q = 0; q < i; ++q) s = fma(-L[i][q], x[q], s);
    x[i] = s * linv[i];
  }
#pragma unroll
  for (int i = NF - 1; i >= 0; --i) {
    double s = x[i];
#pragma unroll
    for (int q = i + 1; q < NF; ++q) s = fma(-L[q][i], x[q], s);
    x[i] = s * linv[i];
  }
}

#ifndef MTG_SOLVE_THREADS
#define MTG_SOLVE_THREADS 128  // threads per CTA (2 per trajectory); 2 CTAs per SM
#endif
template <int HN, int D, bool AOS, int DT = -1>
__global__ void __launch_bounds__(MTG_SOLVE_THREADS, 2) solve_canonical_kernel(const SolveCanonicalParams p) {
  constexpr int N = 2 * HN;
  constexpr int NF = HN - 1;               // free derivatives per interior vertex
  constexpr int SLOTS = NF * NF + NF * D;  // parked G_j and z_j
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ double smem[];
  const int tid = threadIdx.x;
  const int nt = blockDim.x;
  const int pair = (blockIdx.x * nt + tid) >> 1;
  const int side = tid & 1;  // 0: lane A (original orientation), 1: lane B (time-reversed chain)
  const bool active = pair < p.nb;
  const int b = p.b0 + (active ? pair : p.nb - 1);
  const int K = p.K;
  const size_t B = (size_t)p.B;
  const int d = DT >= 0 ? DT : p.derivative;
  uint32_t st = 0;
  // programmatic dependent launch (no-ops for an ordinary launch): let the next solve of the stream start filling
  // the SMs this grid's last, partial wave leaves empty
  if (p.overlap) asm volatile("griddepcontrol.launch_dependents;");

  const size_t rec_pos = (size_t)(K + 1) * D, rec_t = (size_t)K, rec_end = (size_t)2 * NF * D;
  auto pos = [&](int v, int dim) { return p.positions[at<AOS>((size_t)v * D + dim, rec_pos, B, b)]; };
  auto checked_time = [&](double T) {
    if (!(T > 0.0) || !(T < 1.7e308)) {  // LIN_I:296 CHECK_GT(segment_time, 0)
      st |= 1u;
      T = 1.0;
    }
    return T;
  };
  auto seg_time = [&](int i) { return checked_time(p.seg_times[at<AOS>((size_t)i, rec_t, B, b)]); };
  // chain -> original indices
  auto V = [&](int j) { return side ? K - j : j; };
  auto SG = [&](int c) { return side ? K - 1 - c : c; };

  // sign of derivative order m under time reversal (lane B only)
  double sg[HN];
#pragma unroll
  for (int m = 0; m < HN; ++m) sg[m] = (side && (m & 1)) ? -1.0 : 1.0;

  // constraints at the chain start (vertex 0 for A, vertex K for B), chain sign convention
  double sdt[D][HN];
  const bool have_end = (p.end_derivatives != nullptr);
#pragma unroll
  for (int dim = 0; dim < D; ++dim) {
    sdt[dim][0] = pos(V(0), dim);
#pragma unroll
    for (int m = 1; m < HN; ++m)
      sdt[dim][m] = have_end ? sg[m] * p.end_derivatives[at<AOS>((size_t)(side * NF + (m - 1)) * D + dim,
                                                                  rec_end, B, b)]
                             : 0.0;
  }

  double cost_acc = 0.0;

  if (K == 1) {
    if (p.overlap) asm volatile("griddepcontrol.wait;" ::: "memory");
    // fully constrained single segment: lane A writes it from both vertices' constraints
    double other[D][HN];
#pragma unroll
    for (int dim = 0; dim < D; ++dim)
#pragma unroll
      for (int m = 0; m < HN; ++m)  // lane B's constraints arrive in its reversed-time sign convention
        other[dim][m] = ((m & 1) ? -1.0 : 1.0) * __shfl_xor_sync(FULL, sdt[dim][m], 1);
    const double T = seg_time(0);
    cost_acc = emit_segment<HN, D, AOS, DT>(p, 0, b, active && side == 0, T, sdt, other);
    if (active && side == 0) {
      if (p.cost) p.cost[b] = 0.5 * cost_acc;
      if (p.status) p.status[b] = st;
    }
    return;
  }

  const int m_meet = K / 2;                                 // meeting vertex, 1 <= m <= K-1
  const int n_own = side ? (K - 1 - m_meet) : (m_meet - 1);  // vertices this lane eliminates

  // ---------------------------------------------------------------- forward
  double pl[2 * HN - 1], pr[2 * HN - 1];
  double U[NF][NF];  // U_{j-1} on entry of step j (rows: chain vertex j-1, cols: chain vertex j)
  double G[NF][NF];  // G_{j-1}
  double z[D][NF];   // z_{j-1}
  double S[NF][NF];
  double r[D][NF];
  double p_prev[D], p_cur[D], p_next[D];
  segment_powers<HN>(seg_time(SG(0)), d, pr);
#pragma unroll
  for (int dim = 0; dim < D; ++dim) {
    p_cur[dim] = sdt[dim][0];
    p_next[dim] = pos(V(1), dim);
  }
  // inputs of step j are requested during step j-1: the loads (and the division inside
  // segment_powers that waits on them) are off the critical path of the elimination
  double T_ahead = 1.0, p_ahead[D];
#pragma unroll
  for (int dim = 0; dim < D; ++dim) p_ahead[dim] = 0.0;
  if (n_own >= 1) {
    T_ahead = p.seg_times[at<AOS>((size_t)SG(1), rec_t, B, b)];
#pragma unroll
    for (int dim = 0; dim < D; ++dim) p_ahead[dim] = pos(V(2), dim);
  }
#pragma unroll 1
  for (int j = 1;; ++j) {
    const bool own = (j <= n_own);
#pragma unroll
    for (int q = 0; q < 2 * HN - 1; ++q) pl[q] = pr[q];
#pragma unroll
    for (int dim = 0; dim < D; ++dim) {
      p_prev[dim] = p_cur[dim];
      p_cur[dim] = p_next[dim];
    }
    if (own) {
      segment_powers<HN>(checked_time(T_ahead), d, pr);
#pragma unroll
      for (int dim = 0; dim < D; ++dim) p_next[dim] = p_ahead[dim];
      if (j + 1 <= n_own) {
        T_ahead = p.seg_times[at<AOS>((size_t)SG(j + 1), rec_t, B, b)];
#pragma unroll
        for (int dim = 0; dim < D; ++dim) p_ahead[dim] = pos(V(j + 2), dim);
      }
    }
    // contribution of the chain segment on the left of vertex j
#pragma unroll
    for (int k = 0; k < NF; ++k) {
#pragma unroll
      for (int kk = 0; kk <= k; ++kk) S[k][kk] = MTG_H1(HN + 1 + k, HN + 1 + kk) * pl[k + kk + 2];
      const double a_l = MTG_H1(HN + 1 + k, HN) * pl[k + 1];
#pragma unroll
      for (int dim = 0; dim < D; ++dim) r[dim][k] = -a_l * (p_cur[dim] - p_prev[dim]);
    }
    if (d == 0) {  // only a position cost is not translation invariant: H[r][0] + H[r][h] != 0
#pragma unroll
      for (int k = 0; k < NF; ++k) {
        const double a = (MTG_H1(HN + 1 + k, 0) + MTG_H1(HN + 1 + k, HN)) * pl[k + 1];
#pragma unroll
        for (int dim = 0; dim < D; ++dim) r[dim][k] = fma(-a, p_prev[dim], r[dim][k]);
      }
    }
    if (j == 1 && have_end) {
#pragma unroll
      for (int k = 0; k < NF; ++k)
#pragma unroll
        for (int mm = 1; mm < HN; ++mm) {
          const double a = MTG_H1(HN + 1 + k, mm) * pl[k + 1 + mm];
#pragma unroll
          for (int dim = 0; dim < D; ++dim) r[dim][k] = fma(-a, sdt[dim][mm], r[dim][k]);
        }
    }
    if (j > 1) {
#pragma unroll
      for (int k = 0; k < NF; ++k) {
#pragma unroll
        for (int kk = 0; kk <= k; ++kk) {
          double s = S[k][kk];
#pragma unroll
          for (int q = 0; q < NF; ++q) s = fma(-U[q][k], G[q][kk], s);
          S[k][kk] = s;
        }
#pragma unroll
        for (int dim = 0; dim < D; ++dim) {
          double s = r[dim][k];
#pragma unroll
          for (int q = 0; q < NF; ++q) s = fma(-U[q][k], z[dim][q], s);
          r[dim][k] = s;
        }
      }
    }
    if (!own) break;  // j = n_own + 1: S, r hold this lane's half of the meeting vertex
    // contribution of the chain segment on the right of vertex j
#pragma unroll
    for (int k = 0; k < NF; ++k) {
#pragma unroll
      for (int kk = 0; kk <= k; ++kk) S[k][kk] = fma(MTG_H1(1 + k, 1 + kk), pr[k + kk + 2], S[k][kk]);
      const double a_r = MTG_H1(1 + k, HN) * pr[k + 1];
#pragma unroll
      for (int dim = 0; dim < D; ++dim) r[dim][k] = fma(-a_r, p_next[dim] - p_cur[dim], r[dim][k]);
    }
    if (d == 0) {
#pragma unroll
      for (int k = 0; k < NF; ++k) {
        const double a = (MTG_H1(1 + k, 0) + MTG_H1(1 + k, HN)) * pr[k + 1];
#pragma unroll
        for (int dim = 0; dim < D; ++dim) r[dim][k] = fma(-a, p_cur[dim], r[dim][k]);
      }
    }
    double linv[NF];
    chol_lower<NF>(S, linv, st);
#pragma unroll
    for (int dim = 0; dim < D; ++dim) {
#pragma unroll
      for (int i = 0; i < NF; ++i) z[dim][i] = r[dim][i];
      chol_solve<NF>(S, linv, z[dim]);
    }
#pragma unroll
    for (int k = 0; k < NF; ++k)
#pragma unroll
      for (int kk = 0; kk < NF; ++kk) U[k][kk] = MTG_H1(1 + k, HN + 1 + kk) * pr[k + kk + 2];
#pragma unroll
    for (int c = 0; c < NF; ++c) {
      double col[NF];
#pragma unroll
      for (int i = 0; i < NF; ++i) col[i] = U[i][c];
      chol_solve<NF>(S, linv, col);
#pragma unroll
      for (int i = 0; i < NF; ++i) G[i][c] = col[i];
    }
    if (j < n_own) {  // park; the last own vertex stays in registers
      double* slot = smem + (size_t)(j - 1) * SLOTS * nt + tid;
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

  // --------------------------------------------------- meeting vertex m_meet
  // to the original sign convention, add the partner's half, solve (both lanes)
#pragma unroll
  for (int k = 0; k < NF; ++k) {
#pragma unroll
    for (int kk = 0; kk <= k; ++kk) {
      const double mine = S[k][kk] * (sg[k + 1] * sg[kk + 1]);
      S[k][kk] = mine + __shfl_xor_sync(FULL, mine, 1);
    }
#pragma unroll
    for (int dim = 0; dim < D; ++dim) {
      const double mine = r[dim][k] * sg[k + 1];
      r[dim][k] = mine + __shfl_xor_sync(FULL, mine, 1);
    }
  }
  double xe[D][HN];  // values at the chain vertex after the current one, chain sign convention
  {
    double linv[NF];
    chol_lower<NF>(S, linv, st);
#pragma unroll
    for (int dim = 0; dim < D; ++dim) {
      chol_solve<NF>(S, linv, r[dim]);
      xe[dim][0] = p_cur[dim];
#pragma unroll
      for (int i = 0; i < NF; ++i) xe[dim][1 + i] = r[dim][i] * sg[1 + i];
    }
  }
  const size_t rec_free = (size_t)D * (K - 1) * NF;
  // everything above only READ global memory; from here on this grid writes its outputs
  if (p.overlap) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p.free_constraints && active && side == 0) {
#pragma unroll
    for (int dim = 0; dim < D; ++dim)
#pragma unroll
      for (int i = 0; i < NF; ++i)
        p.free_constraints[at<AOS>((size_t)(dim * (K - 1) + (m_meet - 1)) * NF + i, rec_free, B, b)] =
            r[dim][i];
  }

  // ---------------------------------------------------------------- backward
  // fused argmin: the running best cost as of now (any value it has held is a valid filter later); requested here
  // so that the epilogue does not wait for it
  const double best_cost_at_start =
      (p.best_out && tid == 0) ? static_cast<volatile const Best*>(p.best_out)->cost : INFINITY;
  double xs[D][HN];
  double T_back = p.seg_times[at<AOS>((size_t)SG(n_own), rec_t, B, b)];  // requested one step ahead, as above
#pragma unroll 1
  for (int c = n_own; c >= 0; --c) {
    const double T = checked_time(T_back);
    if (c > 0) T_back = p.seg_times[at<AOS>((size_t)SG(c - 1), rec_t, B, b)];
    if (c == 0) {
#pragma unroll
      for (int dim = 0; dim < D; ++dim)
#pragma unroll
        for (int mm = 0; mm < HN; ++mm) xs[dim][mm] = sdt[dim][mm];
    } else {
      if (c < n_own) {
        const double* slot = smem + (size_t)(c - 1) * SLOTS * nt + tid;
#pragma unroll
        for (int i = 0; i < NF; ++i)
#pragma unroll
          for (int q = 0; q < NF; ++q) G[i][q] = slot[(size_t)(i * NF + q) * nt];
#pragma unroll
        for (int dim = 0; dim < D; ++dim)
#pragma unroll
          for (int i = 0; i < NF; ++i) z[dim][i] = slot[(size_t)(NF * NF + dim * NF + i) * nt];
      }
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
        xs[dim][0] = pos(V(c), dim);
#pragma unroll
        for (int i = 0; i < NF; ++i) {
          double s = z[dim][i];
#pragma unroll
          for (int q = 0; q < NF; ++q) s = fma(-G[i][q], xe[dim][1 + q], s);
          xs[dim][1 + i] = s;
        }
      }
      if (p.free_constraints && active) {
        const int v = V(c);
#pragma unroll
        for (int dim = 0; dim < D; ++dim)
#pragma unroll
          for (int i = 0; i < NF; ++i)
            p.free_constraints[at<AOS>((size_t)(dim * (K - 1) + (v - 1)) * NF + i, rec_free, B, b)] =
                xs[dim][1 + i] * sg[1 + i];
      }
    }
    // original orientation / sign: lane B's chain start is the segment's END
    double ds[D][HN], de[D][HN];
#pragma unroll
    for (int dim = 0; dim < D; ++dim)
#pragma unroll
      for (int mm = 0; mm < HN; ++mm) {
        const double a = xs[dim][mm] * sg[mm];
        const double e = xe[dim][mm] * sg[mm];
        ds[dim][mm] = side ? e : a;
        de[dim][mm] = side ? a : e;
      }
    cost_acc += emit_segment<HN, D, AOS, DT>(p, SG(c), b, active, T, ds, de);
#pragma unroll
    for (int dim = 0; dim < D; ++dim)
#pragma unroll
      for (int mm = 0; mm < HN; ++mm) xe[dim][mm] = xs[dim][mm];
  }
  cost_acc += __shfl_xor_sync(FULL, cost_acc, 1);
  st |= __shfl_xor_sync(FULL, st, 1);
  if (active && side == 0) {
    if (p.cost) p.cost[b] = 0.5 * cost_acc;
    if (p.status) p.status[b] = st;
  }
  if (p.best_out) {
    // ---- the sweep's argmin, fused: same total order and the same exclusions (failed solves, NaN) as
    // argmin_kernel; blockDim.x is a multiple of 32 here (the launcher falls back to two launches otherwise).
    // The running best {cost, idx} in *best_out only ever improves, so a CTA whose best is costlier than ANY value
    // the pair has held is done (the value read when the kernel started is such a value: no latency here); the
    // few that might win re-read it and, if still in the race, update it under a lock (whoever waits for the lock
    // keeps re-reading, so a convoy dissolves as soon as a good candidate is in). No per-CTA fence, no partials, no
    // last-CTA fold on the kernel's tail.
    __shared__ double s_bc[32];
    __shared__ long long s_bi[32];
    double c = INFINITY;
    long long i = kInfIdx;
    {
      const double cv = 0.5 * cost_acc;
      if (active && side == 0 && st == 0u && cv == cv) {
        c = cv;
        i = p.best_offset + b;
      }
    }
    warp_reduce(c, i);
    if ((tid & 31) == 0) {
      s_bc[tid >> 5] = c;
      s_bi[tid >> 5] = i;
    }
    __syncthreads();
    if (tid != 0) return;
    for (int q = 1; q < (nt >> 5); ++q)
      if (better(s_bc[q], s_bi[q], c, i)) {
        c = s_bc[q];
        i = s_bi[q];
      }
    if (i == kInfIdx || c > best_cost_at_start) return;
    volatile Best* out = static_cast<volatile Best*>(p.best_out);
    for (;;) {
      if (c > out->cost) return;  // a fresh look before (and while waiting for) the lock: costs only fall
      if (atomicCAS(p.best_lock, 0u, 1u) == 0u) break;
    }
    __threadfence();
    {
      const double bc = out->cost;
      const long long bi = out->idx;
      if (bi < 0 || better(c, i, bc, bi)) {
        out->idx = i;
        out->cost = c;
      }
    }
    __threadfence();
    atomicExch(p.best_lock, 0u);
  }
}

}  // namespace mtg
#endif

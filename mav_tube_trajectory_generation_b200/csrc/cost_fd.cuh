// P9 — finite-difference segment-time perturbation batch.
//
// Replaces, per trajectory and per optimiser evaluation, the caller loop of the reference's
// non-linear layer (include/.../impl/polynomial_optimization_nonlinear_impl.h):
//   getCostAndGradientTime :2495-2584 (central) / getCostAndGradientTimeSimple :2586-2657
//   (forward): for every segment n, T(+-)[n] = T[n] <= 0.1 ? 0.1 : T[n] +- increment_time,
//   updateSegmentTimes (rebuilds Q, A^-1 of ALL K segments, LIN_I:277-304), then
//   getCostAndGradientDerivative :1537-1606: J_d = sum_dim [d_f; d_p]^T R [d_f; d_p] (no 1/2)
//   with the free derivatives d_p HELD FIXED (no re-solve).
//
// R = C^T blockdiag(H_i) C, so J_d = sum_i q_i(T_i) with the per-segment quadratic form
// q_i(T) = T^(1-2d) |W S(T) d_i|^2 (d_i = the segment's 2h endpoint derivatives, H1 = W^T W,
// tables.cpp). A perturbation of T_n changes q_n only: the 1 + 2K full rebuilds of the
// reference collapse to 3 evaluations of one quadratic form per segment:
//   J(+-)[n] = (J - q_n(T_n)) + q_n(T_n +- delta),
//   central:  grad[n] = (q_n(T_n + delta) - q_n(T_n - delta)) / (2 delta)   (no cancellation
//   against the other segments), forward: grad[n] = (q_n(T_n + delta) - q_n(T_n)) / delta.
// One thread per trajectory (SoA: coalesced; AoS: record-strided).
#ifndef MTG_COST_FD_CUH_
#define MTG_COST_FD_CUH_

#include <stdint.h>

#include "device_tables.cuh"
#include "solve_canonical.cuh"  // at<AOS>(), MTG_W

namespace mtg {

struct CostFdParams {
  const double* __restrict__ positions;         // elem (v*D + dim),                  rec (K+1)*D
  const double* __restrict__ end_derivatives;   // elem ((side*(h-1) + m-1)*D + dim), rec 2*(h-1)*D; or nullptr
  const double* __restrict__ seg_times;         // elem i,                            rec K
  const double* __restrict__ free_constraints;  // elem ((dim*(K-1) + v-1)*(h-1) + k-1), rec D*(K-1)*(h-1)
  double* __restrict__ J_nominal;               // [B] or nullptr
  double* __restrict__ J_plus;                  // elem n, rec K; or nullptr
  double* __restrict__ J_minus;                 // elem n, rec K; or nullptr (central only)
  double* __restrict__ grad;                    // elem n, rec K; or nullptr (dJ_d/dT_n)
  uint32_t* __restrict__ status;                // [B] or nullptr
  double increment_time;
  int central;
  int B, b0, nb, K, derivative;
};

// q(T) = T^(1-2d) * sum_dim |W S(T) d|^2 for one segment; ds/de = derivatives 0..HN-1 at its ends
template <int HN, int D>
__device__ __forceinline__ double segment_quadratic(double T, int derivative, const double (&ds)[D][HN],
                                                    const double (&de)[D][HN]) {
  constexpr int N = 2 * HN;
  double tp[HN];
  tp[0] = 1.0;
#pragma unroll
  for (int m = 1; m < HN; ++m) tp[m] = tp[m - 1] * T;
  const int nq = N - derivative;
  double quad = 0.0;
#pragma unroll
  for (int dim = 0; dim < D; ++dim) {
    double dhs[HN], dhe[HN];
    const double dlt = de[dim][0] - ds[dim][0];
#pragma unroll
    for (int m = 1; m < HN; ++m) {
      dhs[m] = tp[m] * ds[dim][m];
      dhe[m] = tp[m] * de[dim][m];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (i < nq) {
        double w = MTG_W(i, HN) * dlt;
        if (derivative == 0) w = fma(MTG_W(i, 0) + MTG_W(i, HN), ds[dim][0], w);
#pragma unroll
        for (int m = 1; m < HN; ++m) {
          w = fma(MTG_W(i, m), dhs[m], w);
          w = fma(MTG_W(i, HN + m), dhe[m], w);
        }
        quad = fma(w, w, quad);
      }
    }
  }
  double s = 1.0;  // T^(1-2d)
  const int e0 = 1 - 2 * derivative;
  if (e0 >= 0) {
    for (int i = 0; i < e0; ++i) s *= T;
  } else {
    const double u = 1.0 / T;
    for (int i = 0; i < -e0; ++i) s *= u;
  }
  return quad * s;
}

template <int HN, int D, bool AOS>
__global__ void __launch_bounds__(128) cost_time_fd_kernel(const CostFdParams p) {
  constexpr int NF = HN - 1;
  const int local = blockIdx.x * blockDim.x + threadIdx.x;
  if (local >= p.nb) return;
  const int b = p.b0 + local;
  const size_t B = (size_t)p.B;
  const int K = p.K;
  const size_t rec_pos = (size_t)(K + 1) * D, rec_end = (size_t)2 * NF * D, rec_free = (size_t)D * (K - 1) * NF;
  uint32_t st = 0;
  // endpoint derivatives of vertex v (fixed position; end constraints or free derivatives)
  auto vertex = [&](int v, double (&d)[D][HN]) {
#pragma unroll
    for (int dim = 0; dim < D; ++dim) {
      d[dim][0] = p.positions[at<AOS>((size_t)v * D + dim, rec_pos, B, b)];
#pragma unroll
      for (int m = 1; m < HN; ++m) {
        double x;
        if (v == 0 || v == K)
          x = p.end_derivatives
                  ? p.end_derivatives[at<AOS>((size_t)((v == 0 ? 0 : 1) * NF + (m - 1)) * D + dim, rec_end, B, b)]
                  : 0.0;
        else
          x = p.free_constraints[at<AOS>((size_t)(dim * (K - 1) + (v - 1)) * NF + (m - 1), rec_free, B, b)];
        d[dim][m] = x;
      }
    }
  };
  const double delta = p.increment_time;
  double ds[D][HN], de[D][HN];
  vertex(0, ds);
  // pass 1: J = sum_i q_i(T_i); the per-segment terms are recomputed in pass 2 (cheaper than parking K values)
  double J = 0.0;
  for (int i = 0; i < K; ++i) {
    double T = p.seg_times[at<AOS>((size_t)i, (size_t)K, B, b)];
    if (!(T > 0.0) || !(T < 1.7e308)) {  // LIN_I:296 CHECK_GT(segment_time, 0)
      st |= 1u;
      T = 1.0;
    }
    vertex(i + 1, de);
    J += segment_quadratic<HN, D>(T, p.derivative, ds, de);
#pragma unroll
    for (int dim = 0; dim < D; ++dim)
#pragma unroll
      for (int m = 0; m < HN; ++m) ds[dim][m] = de[dim][m];
  }
  if (p.J_nominal) p.J_nominal[b] = J;
  if (p.status) p.status[b] = st;
  if (!p.J_plus && !p.J_minus && !p.grad) return;
  vertex(0, ds);
  for (int i = 0; i < K; ++i) {
    double T = p.seg_times[at<AOS>((size_t)i, (size_t)K, B, b)];
    if (!(T > 0.0) || !(T < 1.7e308)) T = 1.0;
    vertex(i + 1, de);
    // NL_I:2527-2530, 2547-2550: the 0.1 s floor replaces the perturbed time, it does not clamp it
    const double Tp = (T <= 0.1) ? 0.1 : T + delta;
    const double q0 = segment_quadratic<HN, D>(T, p.derivative, ds, de);
    const double qp = segment_quadratic<HN, D>(Tp, p.derivative, ds, de);
    const double rest = J - q0;
    const size_t o = at<AOS>((size_t)i, (size_t)K, B, b);
    if (p.J_plus) p.J_plus[o] = rest + qp;
    if (p.central) {
      const double Tm = (T <= 0.1) ? 0.1 : T - delta;
      // 0.1 < T <= increment_time: the shortened time is not positive. The reference aborts there
      // (updateSegmentTimes, LIN_I:296 CHECK_GT(segment_time, 0)); here the item is flagged and gets NaN.
      const bool bad = !(Tm > 0.0);
      if (bad) st |= 1u;
      const double qm = bad ? nan("") : segment_quadratic<HN, D>(Tm, p.derivative, ds, de);
      if (p.J_minus) p.J_minus[o] = rest + qm;
      if (p.grad) p.grad[o] = (qp - qm) / (2.0 * delta);
    } else {
      if (p.grad) p.grad[o] = (qp - q0) / delta;
    }
#pragma unroll
    for (int dim = 0; dim < D; ++dim)
#pragma unroll
      for (int m = 0; m < HN; ++m) ds[dim][m] = de[dim][m];
  }
  if (p.status && st) p.status[b] = st;
}

// setFreeConstraints + updateSegmentsFromCompactConstraints + computeCost (LIN_I:489-498, 254-275,
// 113-130): coefficients and cost of a trajectory whose free derivatives d_p are GIVEN (the
// optimiser-driven path of the non-linear layer, NL_I:1309-1310). One thread per trajectory.
template <int HN, int D, bool AOS>
__global__ void __launch_bounds__(128) coeffs_from_free_kernel(const CostFdParams p, const SolveCanonicalParams sp) {
  constexpr int NF = HN - 1;
  const int local = blockIdx.x * blockDim.x + threadIdx.x;
  if (local >= p.nb) return;
  const int b = p.b0 + local;
  const size_t B = (size_t)p.B;
  const int K = p.K;
  const size_t rec_pos = (size_t)(K + 1) * D, rec_end = (size_t)2 * NF * D, rec_free = (size_t)D * (K - 1) * NF;
  uint32_t st = 0;
  auto vertex = [&](int v, double (&d)[D][HN]) {
#pragma unroll
    for (int dim = 0; dim < D; ++dim) {
      d[dim][0] = p.positions[at<AOS>((size_t)v * D + dim, rec_pos, B, b)];
#pragma unroll
      for (int m = 1; m < HN; ++m) {
        double x;
        if (v == 0 || v == K)
          x = p.end_derivatives
                  ? p.end_derivatives[at<AOS>((size_t)((v == 0 ? 0 : 1) * NF + (m - 1)) * D + dim, rec_end, B, b)]
                  : 0.0;
        else
          x = p.free_constraints[at<AOS>((size_t)(dim * (K - 1) + (v - 1)) * NF + (m - 1), rec_free, B, b)];
        d[dim][m] = x;
      }
    }
  };
  double ds[D][HN], de[D][HN];
  vertex(0, ds);
  double cost = 0.0;
  for (int i = 0; i < K; ++i) {
    double T = p.seg_times[at<AOS>((size_t)i, (size_t)K, B, b)];
    if (!(T > 0.0) || !(T < 1.7e308)) {  // LIN_I:296 CHECK_GT(segment_time, 0)
      st |= 1u;
      T = 1.0;
    }
    vertex(i + 1, de);
    cost += emit_segment<HN, D, AOS>(sp, i, b, true, T, ds, de);
#pragma unroll
    for (int dim = 0; dim < D; ++dim)
#pragma unroll
      for (int m = 0; m < HN; ++m) ds[dim][m] = de[dim][m];
  }
  if (sp.cost) sp.cost[b] = 0.5 * cost;
  if (p.status) p.status[b] = st;
}

// Coefficients and cost from the FULL endpoint derivatives of every vertex (fixed and free entries
// merged by the caller): updateSegmentsFromCompactConstraints + computeCost for any constraint
// pattern (LIN_I:254-275, 113-130). derivs: elem ((v*HN + k)*D + dim), rec (K+1)*HN*D.
template <int HN, int D, bool AOS>
__global__ void __launch_bounds__(128) coeffs_from_derivatives_kernel(const double* __restrict__ derivs,
                                                                      const SolveCanonicalParams sp) {
  const int local = blockIdx.x * blockDim.x + threadIdx.x;
  if (local >= sp.nb) return;
  const int b = sp.b0 + local;
  const size_t B = (size_t)sp.B;
  const int K = sp.K;
  const size_t rec_v = (size_t)(K + 1) * HN * D;
  uint32_t st = 0;
  auto vertex = [&](int v, double (&d)[D][HN]) {
#pragma unroll
    for (int dim = 0; dim < D; ++dim)
#pragma unroll
      for (int m = 0; m < HN; ++m) d[dim][m] = derivs[at<AOS>((size_t)(v * HN + m) * D + dim, rec_v, B, b)];
  };
  double ds[D][HN], de[D][HN];
  vertex(0, ds);
  double cost = 0.0;
  for (int i = 0; i < K; ++i) {
    double T = sp.seg_times[at<AOS>((size_t)i, (size_t)K, B, b)];
    if (!(T > 0.0) || !(T < 1.7e308)) {  // LIN_I:296 CHECK_GT(segment_time, 0)
      st |= 1u;
      T = 1.0;
    }
    vertex(i + 1, de);
    cost += emit_segment<HN, D, AOS>(sp, i, b, true, T, ds, de);
#pragma unroll
    for (int dim = 0; dim < D; ++dim)
#pragma unroll
      for (int m = 0; m < HN; ++m) ds[dim][m] = de[dim][m];
  }
  if (sp.cost) sp.cost[b] = 0.5 * cost;
  if (sp.status) sp.status[b] = st;
}

}  // namespace mtg
#endif

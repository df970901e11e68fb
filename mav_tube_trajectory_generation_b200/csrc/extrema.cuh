// E5 (analytic) / E6 / R1 — extrema of the derivative magnitude |p^(d)(t)| per segment and
// per trajectory.
//
// Replaces (reference): Segment::computeMinMaxMagnitudeCandidateTimes segment.cpp:82-133
// (g = sum_dim conv(delta, delta'), polynomial.cpp:163-181; one dimension: roots of p^(d+1)),
// Polynomial::computeMinMaxCandidates / selectMinMaxCandidatesFromRoots polynomial.cpp:32-83,
// findRootsJenkinsTraub + rpoly_ak1 rpoly_ak1.cpp:57-937, the candidate evaluation
// segment.cpp:135-184 and Trajectory::computeMinMaxMagnitude trajectory.cpp:184-220.
//
// Root finding. The reference runs Jenkins-Traub for ALL complex roots and keeps the real ones
// inside [0, T] (polynomial.cpp:46-60). Only those are ever consumed, so this kernel isolates
// exactly them with the derivative chain: the roots of P^(k+1) split [0, T] into intervals on
// which P^(k) is monotone; every interval whose end values differ in sign holds exactly one
// root of P^(k), refined by a bracketed Newton iteration (bisection when a Newton step leaves
// the bracket). Going from the linear P^(n-1) down to P^(0) = g yields every real root of g in
// [0, T] that is a sign change — i.e. every extremum of the magnitude; roots of even
// multiplicity are inflections of the magnitude and cannot be its minimum or maximum. All loops
// are bounded, there is no data-dependent recursion, and one thread owns one (trajectory,
// segment) problem: the batch supplies the parallelism.
//
// Candidate order and tie rules follow the reference: per segment [t_start, t_end, roots...]
// with std::max / std::min (first wins), across segments strict '>' / '<' (earliest wins).
#ifndef MTG_EXTREMA_CUH_
#define MTG_EXTREMA_CUH_

#include <stdint.h>

#include "device_tables.cuh"
#include "solve_canonical.cuh"  // at<AOS>()

namespace mtg {

constexpr int kMaxG = MTG_BASE_LD;  // 22 coefficients: Polynomial::kMaxConvolutionSize (polynomial.h:48)
constexpr int kRootIters = 96;

struct ExtremaParams {
  const double* __restrict__ coeffs;     // elem ((i*D + dim)*N + j), rec K*D*N
  const double* __restrict__ seg_times;  // elem i, rec K
  double* __restrict__ seg_out;          // scratch [nb][K][4]: min_t, min_v, max_t, max_v (chunk-local)
  uint32_t* __restrict__ seg_status;     // scratch [nb][K]
  double* __restrict__ min_value;        // [B] or nullptr
  double* __restrict__ min_time;         // [B] or nullptr (relative to the segment start, extremum.h:41-42)
  int32_t* __restrict__ min_seg;         // [B] or nullptr
  double* __restrict__ max_value;
  double* __restrict__ max_time;
  int32_t* __restrict__ max_seg;
  double* __restrict__ seg_max_value;    // elem i, rec K; or nullptr: per-segment maxima (candidates of LIN_I:455-487)
  double* __restrict__ seg_max_time;     // elem i, rec K; or nullptr
  uint32_t* __restrict__ status;         // [B] or nullptr
  int B, b0, nb, K, N, D, derivative;
};

// sum_j a[j] t^j, j = 0..deg
__device__ __forceinline__ double horner_n(const double* a, int deg, double t) {
  double r = 0.0;
  for (int j = deg; j >= 0; --j) r = fma(r, t, a[j]);
  return r;
}

template <bool AOS>
__global__ void __launch_bounds__(128) extrema_segment_kernel(const ExtremaParams p) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)p.nb * p.K) return;
  // neighbouring threads take neighbouring trajectories of the SAME segment in both layouts: the
  // work of a root problem depends on the segment (the rest-to-rest end segments carry root
  // clusters), so this keeps a warp's lanes in step; with SoA it also coalesces the loads
  const int local = (int)(gid % p.nb);
  const int seg = (int)(gid / p.nb);
  const int b = p.b0 + local;
  const size_t B = (size_t)p.B;
  const int N = p.N, D = p.D, d = p.derivative, K = p.K;
  const size_t rec_c = (size_t)K * D * N;
  const double T = p.seg_times[at<AOS>((size_t)seg, (size_t)K, B, b)];
  uint32_t st = 0;

  // derivative coefficients delta[dim][j] = B(d, j+d) c[j+d]  (polynomial.h:99-113)
  double delta[4][MTG_TAB_LD];
  const int nd = N - d;  // coefficients of p^(d)
  for (int dim = 0; dim < D; ++dim)
    for (int j = 0; j < MTG_TAB_LD; ++j)
      delta[dim][j] = (j < nd) ? c_base.base[d * MTG_BASE_LD + j + d] *
                                     p.coeffs[at<AOS>((size_t)(seg * D + dim) * N + j + d, rec_c, B, b)]
                               : 0.0;
  // g: polynomial whose real roots in [0, T] are the candidate times
  double g[kMaxG];
  int len;
  if (D > 1) {
    // sum_dim conv(delta, delta'), delta'[j] = (j+1) delta[j+1]   (segment.cpp:93-115)
    len = 2 * nd - 2;
    for (int q = 0; q < kMaxG; ++q) g[q] = 0.0;
    for (int dim = 0; dim < D; ++dim)
      for (int i = 0; i < nd; ++i)
        for (int j = 0; j + 1 < nd; ++j) g[i + j] = fma(delta[dim][i], (double)(j + 1) * delta[dim][j + 1], g[i + j]);
  } else {
    // one dimension: roots of p^(d+1)   (segment.cpp:124-131)
    len = nd - 1;
    for (int q = 0; q < kMaxG; ++q) g[q] = (q < len) ? (double)(q + 1) * delta[0][q + 1] : 0.0;
  }
  // strip zero leading coefficients (findLastNonZeroCoeff, rpoly_ak1.cpp:57-68)
  int n = len - 1;
  while (n >= 0 && !(fabs(g[n]) >= 2.2250738585072014e-308)) --n;

  // ---- real roots of g in [0, T] through the derivative chain
  double ra[kMaxG], rb[kMaxG];  // roots of the previous / current level, ascending
  int na = 0;
  double* prev = ra;
  double* cur = rb;
  const double lo = 0.0, hi = T;
  // Per level two passes, so that the lanes of a warp stay in step: pass 1 only SCANS the partition
  // and collects the brackets with a sign change; pass 2 refines bracket r of every lane together
  // (lanes differ in where their sign changes sit, far less in how many there are).
  double bu[kMaxG], bv[kMaxG], bfu[kMaxG], bfv[kMaxG];
  // roots of the upper levels only PARTITION [0, T] for the level below: 1e-12 T is plenty; the
  // roots of g itself (k = 0) are refined to full precision
  const double tol_upper = 1e-12 * T;
  // coefficients of the current level's polynomial g^(k) and of g^(k+1), built once per level from
  // the table row B(k, .) (polynomial.h:99-113): every evaluation is then a plain Horner sum
  double ca[kMaxG], cb[kMaxG];
  double* pk = ca;
  double* pk1 = cb;
  for (int j = 0; j < kMaxG; ++j) pk[j] = pk1[j] = 0.0;
  if (n >= 1) pk[0] = c_base.base[n * MTG_BASE_LD + n] * g[n];  // g^(n): a constant
  for (int k = n - 1; k >= 0; --k) {
    {
      double* tmpc = pk1;
      pk1 = pk;
      pk = tmpc;
    }
    const int deg = n - k;
    for (int j = 0; j <= deg; ++j) pk[j] = c_base.base[k * MTG_BASE_LD + j + k] * g[j + k];
    int nb = 0;
    double u = lo, fu = horner_n(pk, deg, u);
    for (int q = 0; q <= na; ++q) {
      const double v = (q < na) ? prev[q] : hi;
      if (!(v > u)) continue;
      const double fv = horner_n(pk, deg, v);
      if (fu == 0.0) {  // a root exactly on a partition point
        if (nb == 0 || bu[nb - 1] != u || bfu[nb - 1] != 0.0) {
          bu[nb] = u;
          bv[nb] = u;
          bfu[nb] = 0.0;
          bfv[nb] = 0.0;
          ++nb;
        }
      } else if (fv != 0.0 && ((fu < 0.0) != (fv < 0.0))) {
        bu[nb] = u;
        bv[nb] = v;
        bfu[nb] = fu;
        bfv[nb] = fv;
        ++nb;
      }
      u = v;
      fu = fv;
    }
    if (fu == 0.0 && (nb == 0 || bu[nb - 1] != u || bfu[nb - 1] != 0.0)) {  // root exactly at t_end
      bu[nb] = u;
      bv[nb] = u;
      bfu[nb] = 0.0;
      bfv[nb] = 0.0;
      ++nb;
    }
    // Refinement, flattened: ONE loop whose trip is one bracketed-Newton iteration of whichever bracket the
    // lane is working on (the iterates of refine_root, unchanged). The lanes of a warp then wait for the
    // largest SUM of iterations instead of the sum over brackets of the largest iteration count.
    {
      const double tol = k ? tol_upper : 0.0;
      int r = 0, it = 0;
      bool have = false;
      double a = 0.0, bb = 0.0, fa = 0.0, t = 0.0;
      for (;;) {
        if (!have) {
          while (r < nb && bfu[r] == 0.0) {  // a root exactly on a partition point
            cur[r] = bu[r];
            ++r;
          }
          if (r >= nb) break;
          a = bu[r];
          bb = bv[r];
          fa = bfu[r];
          const double fb = bfv[r];
          t = a - fa * ((bb - a) / (fb - fa));  // first iterate: the secant point (the midpoint if it degenerates)
          if (!(t > a && t < bb)) t = 0.5 * (a + bb);
          it = 0;
          have = true;
        }
        // value and slope in one sweep (two independent FMA chains)
        double ft = pk[deg], dft = 0.0;
        for (int j = deg - 1; j >= 0; --j) {
          dft = fma(dft, t, pk1[j]);
          ft = fma(ft, t, pk[j]);
        }
        bool fin = ft == 0.0;
        double res = t;
        if (!fin) {
          if ((ft < 0.0) == (fa < 0.0)) {
            a = t;
            fa = ft;
          } else {
            bb = t;
          }
          const double width = bb - a;
          if (!(width > fmax(tol, 4.5e-16 * fmax(fabs(a), fabs(bb))))) {
            fin = true;
          } else {
            double tn = t - ft / dft;
            if (!(tn > a && tn < bb)) tn = 0.5 * (a + bb);
            if (!(fabs(tn - t) > tol)) {
              fin = true;
              res = tn;
            } else {
              t = tn;
              if (++it >= kRootIters) {
                st |= 16u;  // MTG_ST_NO_CONVERGENCE (reference: rpoly prints and returns partial roots, RPOLY_C:372-377)
                fin = true;
                res = t;
              }
            }
          }
        }
        if (fin) {
          cur[r] = res;
          ++r;
          have = false;
        }
      }
    }
    double* tmp = prev;
    prev = cur;
    cur = tmp;
    na = nb;
  }
  if (n < 1) na = 0;  // constant polynomial: no roots (rpoly_ak1.cpp:76-80)

  // ---- candidates [t_start, t_end, roots...]: |p^(d)(t)| = sqrt(sum_dim evaluate(t, d)^2)  (segment.cpp:135-158)
  double mn_v = 1.7976931348623157e308, mx_v = -1.7976931348623157e308, mn_t = 0.0, mx_t = 0.0;
  for (int q = 0; q < na + 2; ++q) {
    const double t = (q == 0) ? lo : (q == 1) ? hi : prev[q - 2];
    if (t < lo || t > hi) continue;  // also drops NaN times
    double m2 = 0.0;
    for (int dim = 0; dim < D; ++dim) {
      double r = 0.0;
      for (int j = nd - 1; j >= 0; --j) r = fma(r, t, delta[dim][j]);
      m2 = fma(r, r, m2);
    }
    const double m = sqrt(m2);
    if (mx_v < m) {  // std::max keeps the first on ties
      mx_v = m;
      mx_t = t;
    }
    if (m < mn_v) {
      mn_v = m;
      mn_t = t;
    }
  }
  double* o = p.seg_out + ((size_t)local * K + seg) * 4;
  o[0] = mn_t;
  o[1] = mn_v;
  o[2] = mx_t;
  o[3] = mx_v;
  p.seg_status[(size_t)local * K + seg] = st;
}

// Trajectory::computeMinMaxMagnitude, trajectory.cpp:184-220: strict comparisons in segment order
template <bool AOS>
__global__ void __launch_bounds__(256) extrema_reduce_kernel(const ExtremaParams p) {
  const int local = blockIdx.x * blockDim.x + threadIdx.x;
  if (local >= p.nb) return;
  const int b = p.b0 + local;
  const int K = p.K;
  double mn_v = 1.7976931348623157e308, mx_v = -1.7976931348623157e308, mn_t = 0.0, mx_t = 0.0;
  int mn_s = 0, mx_s = 0;
  uint32_t st = 0;
  for (int s = 0; s < K; ++s) {
    const double* o = p.seg_out + ((size_t)local * K + s) * 4;
    const double a_t = o[0], a_v = o[1], z_t = o[2], z_v = o[3];
    st |= p.seg_status[(size_t)local * K + s];
    if (a_v < mn_v) {
      mn_v = a_v;
      mn_t = a_t;
      mn_s = s;
    }
    if (z_v > mx_v) {
      mx_v = z_v;
      mx_t = z_t;
      mx_s = s;
    }
    if (p.seg_max_value) p.seg_max_value[at<AOS>((size_t)s, (size_t)K, (size_t)p.B, (size_t)b)] = z_v;
    if (p.seg_max_time) p.seg_max_time[at<AOS>((size_t)s, (size_t)K, (size_t)p.B, (size_t)b)] = z_t;
  }
  if (p.min_value) p.min_value[b] = mn_v;
  if (p.min_time) p.min_time[b] = mn_t;
  if (p.min_seg) p.min_seg[b] = mn_s;
  if (p.max_value) p.max_value[b] = mx_v;
  if (p.max_time) p.max_time[b] = mx_t;
  if (p.max_seg) p.max_seg[b] = mx_s;
  if (p.status) p.status[b] = st;
}

}  // namespace mtg
#endif

// Sampled evaluation and feasibility sweep — one THREAD per trajectory marching
// over its samples, so that the reference's SERIAL sampling recurrence
//   acc += dt;  tau += dt;  tau -= T_i on crossing (strict >)       trajectory.cpp:74-134
// is replayed bit-exactly (segment index, sample count and sampling time of every
// sample are integers / IEEE sums that must match the reference), while a warp
// writes one sample row of 32 neighbouring trajectories per store (SoA, fully
// coalesced 256-B rows). The kernels are HBM-write bound (24 B per sample).
//
// Replaces (reference): Polynomial::evaluate polynomial.h:136-149, Segment::evaluate
// segment.cpp:51-58, Trajectory::evaluate / evaluateRange trajectory.cpp:41-134,
// the sampled limit check test_utils.h:43-54 / NL_I:2686-2733, and the sampled form
// of the tube geometry polynomial_optimization_qcqp_impl.h:357-474.
#ifndef MTG_EVAL_CUH_
#define MTG_EVAL_CUH_

#include <stdint.h>

#include "device_tables.cuh"
#include "solve_canonical.cuh"  // at<AOS>()

namespace mtg {

struct EvalParams {
  const double* __restrict__ coeffs;      // elem ((i*D + dim)*N + j), rec K*D*N
  const double* __restrict__ seg_times;   // elem i, rec K
  const double* __restrict__ t_start;     // [B]
  const double* __restrict__ t_end;       // [B]
  const double* __restrict__ dt;          // [B]
  double* __restrict__ samples;           // elem (n*D + dim), rec max_samples*D ; or nullptr
  double* __restrict__ sampling_times;    // elem n, rec max_samples ; or nullptr
  int32_t* __restrict__ segment_idx;      // elem n, rec max_samples ; or nullptr
  int32_t* __restrict__ n_samples;        // [B] or nullptr
  uint32_t* __restrict__ status;          // [B] or nullptr
  // feasibility sweep only
  const double* __restrict__ positions;   // vertices, elem (v*3 + dim), rec (K+1)*3 ; or nullptr
  const double* __restrict__ radii;       // elem (i*2 + {0,1}), rec K*2 ; or nullptr
  uint8_t* __restrict__ flags;            // elem n, rec max_samples ; or nullptr
  double* __restrict__ max_v;             // [B] or nullptr
  double* __restrict__ max_a;             // [B] or nullptr
  uint8_t* __restrict__ feasible;         // [B] or nullptr
  double v_max, a_max;
  double v2_lim, a2_lim;  // largest doubles whose sqrt is <= v_max / a_max (time-major sweep)
  int B, b0, nb, K, N;  // N = coefficients actually stored (<= NT of the kernel)
  int derivative;
  int max_samples;
  int vec_ok;
};

// coefficient j of polynomial (seg, dim) of trajectory b
template <bool AOS>
__device__ __forceinline__ double load_coeff(const EvalParams& p, int seg, int dim, int D, int j, int b) {
  return p.coeffs[at<AOS>((size_t)(seg * D + dim) * p.N + j, (size_t)p.K * D * p.N, (size_t)p.B, (size_t)b)];
}

// Loads segment `seg` into registers as the coefficients of its `derivative`-th
// derivative: c[dim][i] = B(derivative, i + derivative) * coeff[i + derivative]
// (polynomial.h:99-113), zero above. Horner over all NT slots then reproduces
// polynomial.h:136-149 (leading zero slots are exact no-ops).
template <int NT, int D, bool AOS>
__device__ __forceinline__ void load_segment(const EvalParams& p, int seg, int b, int derivative,
                                             double (&c)[D][NT]) {
#pragma unroll
  for (int dim = 0; dim < D; ++dim) {
    double raw[NT];
    if (AOS && p.vec_ok && p.N == NT) {
      const double2* src = reinterpret_cast<const double2*>(
          p.coeffs + (size_t)b * ((size_t)p.K * D * NT) + (size_t)(seg * D + dim) * NT);
#pragma unroll
      for (int j = 0; j < NT; j += 2) {
        const double2 v = src[j / 2];
        raw[j] = v.x;
        raw[j + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NT; ++j) raw[j] = (j < p.N) ? load_coeff<AOS>(p, seg, dim, D, j, b) : 0.0;
    }
    if (derivative == 0) {
#pragma unroll
      for (int j = 0; j < NT; ++j) c[dim][j] = raw[j];
    } else {
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < NT; ++j)  // select raw[i + derivative] without dynamic register indexing
          if (j == i + derivative) v = raw[j] * c_base.base[derivative * MTG_BASE_LD + j];
        c[dim][i] = v;
      }
    }
  }
}

template <int NT>
__device__ __forceinline__ double horner(const double (&c)[NT], double t) {
  double r = c[NT - 1];
#pragma unroll
  for (int j = NT - 2; j >= 0; --j) r = fma(r, t, c[j]);
  return r;
}

// ------------------------------------------------------------ evaluateRange
template <int NT, int D, bool AOS>
__global__ void __launch_bounds__(128) eval_range_kernel(const EvalParams p) {
  const int local = blockIdx.x * blockDim.x + threadIdx.x;
  if (local >= p.nb) return;
  const int b = p.b0 + local;
  const size_t B = (size_t)p.B;
  const int K = p.K;
  const size_t rec_t = (size_t)K;
  const size_t rec_s = (size_t)p.max_samples;
  uint32_t st = 0;
  int n = 0;
  const double t0 = p.t_start[b], t1 = p.t_end[b], dt = p.dt[b];

  // trajectory.cpp:88-111: first segment whose accumulated end time exceeds t_start
  double acc = 0.0;
  int i = 0;
  double Ti = 0.0;
  for (i = 0; i < K; ++i) {
    Ti = p.seg_times[at<AOS>((size_t)i, rec_t, B, b)];
    acc += Ti;
    if (acc > t0) break;
  }
  if (t0 > acc || i >= K || !(dt > 0.0)) {
    // out of range (reference: LOG(ERROR), empty result; t_start == max_time is UB there)
    st |= 4u;
  } else {
    acc -= Ti;
    double tau = t0 - acc;
    double c[D][NT];
    load_segment<NT, D, AOS>(p, i, b, p.derivative, c);
    while (acc < t1) {
      bool off_end = false;
      while (tau > Ti) {
        tau = tau - Ti;
        ++i;
        if (i >= K) {
          off_end = true;
          break;
        }
        Ti = p.seg_times[at<AOS>((size_t)i, rec_t, B, b)];
        load_segment<NT, D, AOS>(p, i, b, p.derivative, c);
      }
      if (off_end) break;
      if (n >= p.max_samples) {
        st |= 8u;
        break;
      }
      if (p.samples) {
#pragma unroll
        for (int dim = 0; dim < D; ++dim)
          p.samples[at<AOS>((size_t)n * D + dim, rec_s * D, B, b)] = horner<NT>(c[dim], tau);
      }
      if (p.sampling_times) p.sampling_times[at<AOS>((size_t)n, rec_s, B, b)] = acc;
      if (p.segment_idx) p.segment_idx[at<AOS>((size_t)n, rec_s, B, b)] = i;
      tau += dt;
      acc += dt;
      ++n;
    }
  }
  if (p.n_samples) p.n_samples[b] = n;
  if (p.status) p.status[b] = st;
}

// ------------------------------------------------------- Trajectory::evaluate
struct EvalAtParams {
  const double* __restrict__ coeffs;
  const double* __restrict__ seg_times;
  const double* __restrict__ t;          // elem m, rec M
  double* __restrict__ out;              // elem (m*D + dim), rec M*D
  int32_t* __restrict__ segment_idx;     // elem m, rec M ; or nullptr (-1 = out of range)
  uint32_t* __restrict__ status;         // [B] or nullptr (OR of all queries; must be pre-zeroed)
  int B, b0, nb, K, N, D, M, derivative;
};

template <bool AOS>
__global__ void __launch_bounds__(256) eval_at_kernel(const EvalAtParams p) {
  // thread = (query m, trajectory b); SoA: b fastest so that a warp reads neighbouring records
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)p.nb * p.M;
  if (gid >= total) return;
  const int local = AOS ? (int)(gid / p.M) : (int)(gid % p.nb);
  const int m = AOS ? (int)(gid % p.M) : (int)(gid / p.nb);
  const int b = p.b0 + local;
  const size_t B = (size_t)p.B;
  const int K = p.K, N = p.N, D = p.D;
  const double t = p.t[at<AOS>((size_t)m, (size_t)p.M, B, b)];
  // trajectory.cpp:41-72
  double acc = 0.0, Ti = 0.0;
  int i = 0;
  for (i = 0; i < K; ++i) {
    Ti = p.seg_times[at<AOS>((size_t)i, (size_t)K, B, b)];
    acc += Ti;
    if (acc > t) break;
  }
  const bool out_of_range = (t > acc) || !(t == t);
  if (i >= K) {
    i = K - 1;  // t == max time: last segment
    Ti = p.seg_times[at<AOS>((size_t)i, (size_t)K, B, b)];
  }
  acc -= Ti;
  const double tau = t - acc;
  const size_t rec_c = (size_t)K * D * N;
  for (int dim = 0; dim < D; ++dim) {
    double r = 0.0;
    if (!out_of_range && p.derivative < N) {
      const size_t e0 = (size_t)(i * D + dim) * N;
      r = c_base.base[p.derivative * MTG_BASE_LD + (N - 1)] * p.coeffs[at<AOS>(e0 + N - 1, rec_c, B, b)];
      for (int j = N - 2; j >= p.derivative; --j)
        r = fma(r, tau, c_base.base[p.derivative * MTG_BASE_LD + j] * p.coeffs[at<AOS>(e0 + j, rec_c, B, b)]);
    }
    p.out[at<AOS>((size_t)m * D + dim, (size_t)p.M * D, B, b)] = r;  // zeros when out of range (TRAJ_C:58-61)
  }
  if (p.segment_idx) p.segment_idx[at<AOS>((size_t)m, (size_t)p.M, B, b)] = out_of_range ? -1 : i;
  if (out_of_range && p.status) atomicOr(&p.status[b], 4u);
}

// --------------------------------------------------------- feasibility sweep
// Same sampling recurrence; evaluates position, velocity and acceleration of every
// sample with one fused triple-Horner (p, p', p''/2 by repeated synthetic division),
// writes per-sample flags and per-trajectory maxima.
//   flags bit0: |v| <= v_max   bit1: |a| <= a_max   bit2: inside the tube of the segment
// Tube of segment i (sampled restatement of QC_I:369-474, D = 3 only):
//   n = (v_{i+1}-v_i)/|.|, A = I - n n^T (entries |.|<1e-6 zeroed), b = -A v_i (same zeroing)
//   inside <=> |A x + b|^2 <= r_i^2  and  n.(x - p_start) >= 0  and  n.(x - p_end) <= 0
//   p_start = v_i - n r_start (r_start = radii[0].first if i == 0 else radii[i-1].second),
//   p_end = v_{i+1} + n radii[i].second.
struct TubeSeg {
  double A[6];  // symmetric: xx xy xz yy yz zz
  double bvec[3];
  double n[3];
  double cs, ce;  // n.p_start, n.p_end
  double r2;
};

template <bool AOS>
__device__ __forceinline__ void load_tube(const EvalParams& p, int i, int b, TubeSeg& t) {
  const size_t B = (size_t)p.B;
  const size_t rec_p = (size_t)(p.K + 1) * 3, rec_r = (size_t)p.K * 2;
  double s[3], e[3], v[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    s[k] = p.positions[at<AOS>((size_t)i * 3 + k, rec_p, B, b)];
    e[k] = p.positions[at<AOS>((size_t)(i + 1) * 3 + k, rec_p, B, b)];
    v[k] = e[k] - s[k];
  }
  // no FMA contraction here: the 1e-6 zeroing thresholds make these values discrete
  const double nrm = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(v[0], v[0]), __dmul_rn(v[1], v[1])), __dmul_rn(v[2], v[2])));
#pragma unroll
  for (int k = 0; k < 3; ++k) v[k] = v[k] / nrm;
  const double nx = v[0], ny = v[1], nz = v[2];
  const double px = s[0], py = s[1], pz = s[2];
  auto z6 = [](double x) { return (x > -0.000001 && x < 0.000001) ? 0.0 : x; };
  const double nxx = __dmul_rn(nx, nx), nyy = __dmul_rn(ny, ny), nzz = __dmul_rn(nz, nz);
  const double nxy = __dmul_rn(nx, ny), nxz = __dmul_rn(nx, nz), nyz = __dmul_rn(ny, nz);
  t.A[0] = z6(1 - nxx);
  t.A[1] = z6(-nxy);
  t.A[2] = z6(-nxz);
  t.A[3] = z6(1 - nyy);
  t.A[4] = z6(-nyz);
  t.A[5] = z6(1 - nzz);
  t.bvec[0] = z6(__dadd_rn(__dadd_rn(__dmul_rn(nxx - 1, px), __dmul_rn(nxy, py)), __dmul_rn(nxz, pz)));
  t.bvec[1] = z6(__dadd_rn(__dadd_rn(__dmul_rn(nxy, px), __dmul_rn(nyy - 1, py)), __dmul_rn(nyz, pz)));
  t.bvec[2] = z6(__dadd_rn(__dadd_rn(__dmul_rn(nxz, px), __dmul_rn(nyz, py)), __dmul_rn(nzz - 1, pz)));
  const double r_tube = p.radii[at<AOS>((size_t)i * 2, rec_r, B, b)];
  const double r_end = p.radii[at<AOS>((size_t)i * 2 + 1, rec_r, B, b)];
  const double r_start = (i == 0) ? p.radii[at<AOS>((size_t)0, rec_r, B, b)]
                                  : p.radii[at<AOS>((size_t)(i - 1) * 2 + 1, rec_r, B, b)];
  t.cs = 0.0;
  t.ce = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    t.n[k] = v[k];
    t.cs += v[k] * (s[k] - v[k] * r_start);
    t.ce += v[k] * (e[k] + v[k] * r_end);
  }
  t.r2 = r_tube * r_tube;
}

__device__ __forceinline__ bool in_tube(const TubeSeg& t, const double (&x)[3]) {
  const double y0 = t.A[0] * x[0] + t.A[1] * x[1] + t.A[2] * x[2] + t.bvec[0];
  const double y1 = t.A[1] * x[0] + t.A[3] * x[1] + t.A[4] * x[2] + t.bvec[1];
  const double y2 = t.A[2] * x[0] + t.A[4] * x[1] + t.A[5] * x[2] + t.bvec[2];
  const double q = y0 * y0 + y1 * y1 + y2 * y2;
  const double along = t.n[0] * x[0] + t.n[1] * x[1] + t.n[2] * x[2];
  return (q <= t.r2) && (along >= t.cs) && (along <= t.ce);
}

template <int NT, int D, bool AOS>
__global__ void __launch_bounds__(128) feasibility_kernel(const EvalParams p) {
  const int local = blockIdx.x * blockDim.x + threadIdx.x;
  if (local >= p.nb) return;
  const int b = p.b0 + local;
  const size_t B = (size_t)p.B;
  const int K = p.K;
  const size_t rec_t = (size_t)K;
  const size_t rec_s = (size_t)p.max_samples;
  const bool tube = (D == 3) && p.radii != nullptr && p.positions != nullptr;
  uint32_t st = 0;
  int n = 0;
  double mv = 0.0, ma = 0.0;
  unsigned all_bits = 7u;
  const double t0 = p.t_start[b], t1 = p.t_end[b], dt = p.dt[b];
  double acc = 0.0, Ti = 0.0;
  int i = 0;
  for (i = 0; i < K; ++i) {
    Ti = p.seg_times[at<AOS>((size_t)i, rec_t, B, b)];
    acc += Ti;
    if (acc > t0) break;
  }
  if (t0 > acc || i >= K || !(dt > 0.0)) {
    st |= 4u;
  } else {
    acc -= Ti;
    double tau = t0 - acc;
    double c[D][NT];
    TubeSeg ts;
    load_segment<NT, D, AOS>(p, i, b, 0, c);
    if (tube) load_tube<AOS>(p, i, b, ts);
    while (acc < t1) {
      bool off_end = false;
      while (tau > Ti) {
        tau = tau - Ti;
        ++i;
        if (i >= K) {
          off_end = true;
          break;
        }
        Ti = p.seg_times[at<AOS>((size_t)i, rec_t, B, b)];
        load_segment<NT, D, AOS>(p, i, b, 0, c);
        if (tube) load_tube<AOS>(p, i, b, ts);
      }
      if (off_end) break;
      if (n >= p.max_samples) {
        st |= 8u;
        break;
      }
      double x[D], v2 = 0.0, a2 = 0.0;
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
        double p0 = c[dim][NT - 1], p1 = 0.0, p2 = 0.0;
#pragma unroll
        for (int j = NT - 2; j >= 0; --j) {
          p2 = fma(p2, tau, p1);
          p1 = fma(p1, tau, p0);
          p0 = fma(p0, tau, c[dim][j]);
        }
        x[dim] = p0;
        v2 = fma(p1, p1, v2);
        a2 = fma(2.0 * p2, 2.0 * p2, a2);
      }
      const double nv = sqrt(v2), na = sqrt(a2);
      mv = fmax(mv, nv);
      ma = fmax(ma, na);
      unsigned f = (nv <= p.v_max ? 1u : 0u) | (na <= p.a_max ? 2u : 0u);
      if (D == 3) {
        if (tube) {
          const double x3[3] = {x[0], x[D > 1 ? 1 : 0], x[D > 2 ? 2 : 0]};
          f |= in_tube(ts, x3) ? 4u : 0u;
        } else {
          f |= 4u;
        }
      } else {
        f |= 4u;
      }
      all_bits &= f;
      if (p.samples) {
#pragma unroll
        for (int dim = 0; dim < D; ++dim) p.samples[at<AOS>((size_t)n * D + dim, rec_s * D, B, b)] = x[dim];
      }
      if (p.flags) p.flags[at<AOS>((size_t)n, rec_s, B, b)] = (uint8_t)f;
      tau += dt;
      acc += dt;
      ++n;
    }
  }
  if (p.n_samples) p.n_samples[b] = n;
  if (p.max_v) p.max_v[b] = mv;
  if (p.max_a) p.max_a[b] = ma;
  if (p.feasible) p.feasible[b] = (uint8_t)((all_bits == 7u && st == 0) ? 1 : 0);
  if (p.status) p.status[b] = st;
}

// max time of each trajectory, summed in segment order like Trajectory::addSegments
// (trajectory.h:63-72): max_time_ += segment.getTime()
template <bool AOS>
__global__ void max_time_kernel(const double* __restrict__ seg_times, double* __restrict__ out, int B, int b0,
                                int nb, int K) {
  const int local = blockIdx.x * blockDim.x + threadIdx.x;
  if (local >= nb) return;
  const int b = b0 + local;
  double acc = 0.0;
  for (int i = 0; i < K; ++i) acc += seg_times[at<AOS>((size_t)i, (size_t)K, (size_t)B, (size_t)b)];
  out[b] = acc;
}

}  // namespace mtg
#endif

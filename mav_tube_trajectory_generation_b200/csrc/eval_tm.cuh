// Time-major sampled evaluation / feasibility sweep ("tm" kernels) — the path for
// trajectory-contiguous sample outputs (the reference's own order: evaluateRange
// fills one std::vector<VectorXd> per trajectory, trajectory.cpp:74-134).
//
// One WARP owns 32 trajectories and alternates two phases over chunks of 32 samples:
//
//  phase 1  lane = trajectory. Every lane replays the reference's SERIAL sampling
//           recurrence (acc += dt; tau += dt; tau -= T_i on a strict '>' crossing)
//           for up to 32 samples of its trajectory and parks (tau, segment[, acc])
//           in shared memory. Only DADDs and compares: the part of the algorithm
//           that cannot be parallelised over samples costs ~3 fp64 ops per sample.
//  phase 2  lane = sample. For each of the 32 trajectories in turn the warp
//           evaluates 32 consecutive samples at once: broadcast 16-byte loads of the
//           segment's coefficients, Horner (fused multiply-adds), and the D*32
//           doubles of the sample rows leave the warp as whole, consecutive 256-byte
//           stores (staged through shared memory). No divergence: a lane whose
//           sample lies in the next segment simply reads another address.
//
// The warp is self-contained (only __syncwarp), so occupancy is a pure launch knob.
// Replaces (reference): Polynomial::evaluate polynomial.h:136-149, Segment::evaluate
// segment.cpp:51-58, Trajectory::evaluateRange trajectory.cpp:74-134, the sampled
// limit check test_utils.h:43-54 / NL_I:2686-2733 and the sampled form of the tube
// geometry polynomial_optimization_qcqp_impl.h:357-474.
#ifndef MTG_EVAL_TM_CUH_
#define MTG_EVAL_TM_CUH_

#include <stdint.h>

#include "eval.cuh"

namespace mtg {

constexpr int kTmChunk = 32;              // samples per trajectory per chunk (= lanes)
constexpr int kTmTauLd = kTmChunk + 1;    // odd stride in 8-byte words: conflict-free lane-major stores
constexpr int kTubeGeomLd = 16;           // doubles per (trajectory, segment) tube record (one 128-B line)

__device__ __forceinline__ double2 ldg_nc2(const double2* p) {
  double2 v;
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_group1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// max over the warp of non-negative doubles (bit patterns order like the values)
__device__ __forceinline__ double warp_max_nonneg(double x) {
  const unsigned hi = (unsigned)__double2hiint(x), lo = (unsigned)__double2loint(x);
  const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
  return __hiloint2double((int)mhi, (int)mlo);
}

// trajectory.cpp:88-111: first segment whose accumulated end time exceeds t_start; on
// success acc is the START time of that segment, computed as (sum_{j<=i} T_j) - T_i like
// the reference. false: t_start out of range (reference: LOG(ERROR) + empty result;
// t_start == max_time indexes segments_[K] there and is an error here) or dt <= 0.
template <bool AOS>
__device__ __forceinline__ bool locate_start(const EvalParams& p, int b, double t0, double dt, int& i,
                                             double& acc) {
  acc = 0.0;
  double Ti = 0.0;
  for (i = 0; i < p.K; ++i) {
    Ti = p.seg_times[at<AOS>((size_t)i, (size_t)p.K, (size_t)p.B, (size_t)b)];
    acc += Ti;
    if (acc > t0) break;
  }
  if (t0 > acc || i >= p.K || !(dt > 0.0)) return false;
  acc -= Ti;
  return true;
}

// ---------------------------------------------------------------- tube setup
// One thread per (trajectory, segment): the constants of the sampled tube predicate
// (QC_I:369-474, see eval.cuh) as one 128-byte record: A[6] b[3] n[3] cs ce r2 pad.
template <bool AOS>
__global__ void __launch_bounds__(256) tube_setup_kernel(const EvalParams p, double* __restrict__ geom) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)p.nb * p.K) return;
  const int local = AOS ? (int)(gid / p.K) : (int)(gid % p.nb);
  const int seg = AOS ? (int)(gid % p.K) : (int)(gid / p.nb);
  TubeSeg t;
  load_tube<AOS>(p, seg, p.b0 + local, t);
  double2* o = reinterpret_cast<double2*>(geom + ((size_t)local * p.K + seg) * kTubeGeomLd);
  o[0] = make_double2(t.A[0], t.A[1]);
  o[1] = make_double2(t.A[2], t.A[3]);
  o[2] = make_double2(t.A[4], t.A[5]);
  o[3] = make_double2(t.bvec[0], t.bvec[1]);
  o[4] = make_double2(t.bvec[2], t.n[0]);
  o[5] = make_double2(t.n[1], t.n[2]);
  o[6] = make_double2(t.cs, t.ce);
  o[7] = make_double2(t.r2, 0.0);
}

// ------------------------------------------------------------------ the sweep
// FEAS = false: samples of derivative p.derivative (+ sampling_times, segment_idx).
// FEAS = true : position samples (optional) + v/a/tube flags + per-trajectory maxima.
// Inputs (coeffs, seg_times, positions, radii) in AOS_IN layout; every per-sample
// output is trajectory-contiguous: x[b * max_samples * width + n * width + ...].
//
// Shared memory per warp (TmLayout): tau[32][33] | output staging rows | info[32] |
// acc[32][33] (only when sampling_times is requested) | per-trajectory SEGMENT SLOTS.
// A slot pair holds the records {coefficients, tube constants, duration} of the two
// segments a chunk may touch: segment s lives in slot s & 1, is fetched ONCE with
// cp.async (LDGSTS) by the trajectory's lane, and the next segment is prefetched right
// after a chunk has been evaluated, a few chunks before it is needed, so neither phase
// waits on DRAM in steady state (one commit group per chunk; wait_group 1 retires all
// but the newest). A chunk stops early if a third segment would start (tiny segments
// or large dt).
struct TmLayout {
  int slot_bytes;   // one segment: D*NT coefficients (+ tube record) + {T, pad}; multiple of 16
  int traj_bytes;   // 2 slots + 16 B pad (odd multiple of 16 B: conflict-free lane-major cp.async)
  int t_off;        // byte offset of the segment duration inside a slot
  int off_stage, off_info, off_acc, off_slots, per_warp;
};
constexpr int kTmU = 2;  // trajectories evaluated together in phase 2 (independent FMA chains)
__host__ __device__ inline TmLayout tm_layout(int D, int NT, bool want_acc, bool tube, bool slots) {
  TmLayout L;
  L.t_off = D * NT * 8 + (tube ? kTubeGeomLd * 8 : 0);
  L.slot_bytes = L.t_off + 16;
  L.traj_bytes = 2 * L.slot_bytes + 16;
  L.off_stage = 32 * kTmTauLd * 8;
  L.off_info = L.off_stage + kTmU * 32 * D * 8;
  L.off_acc = L.off_info + 32 * 16;
  L.off_slots = L.off_acc + (want_acc ? 32 * kTmTauLd * 8 : 0);
  L.per_warp = L.off_slots + (slots ? 32 * L.traj_bytes : 0);
  L.per_warp = (L.per_warp + 15) & ~15;
  return L;
}

enum TmMode { TM_POSITION = 0, TM_DERIVATIVE = 1, TM_FEAS = 2, TM_FEAS_TUBE = 3 };

// Requirements (checked by the launcher, which otherwise falls back to the one-thread-per-
// trajectory kernels of eval.cuh): AoS layout, N == NT, coeffs/seg_times/geom 16-/8-byte aligned.
template <int NT, int D, int MODE>
__global__ void __launch_bounds__(64, 4) eval_tm_kernel(const EvalParams p, const double* __restrict__ geom) {
  constexpr bool FEAS = MODE >= TM_FEAS;
  constexpr bool AOS_IN = true;
  constexpr bool fast_c = true;  // segment records staged in shared memory
  constexpr bool tube = MODE == TM_FEAS_TUBE;
  static_assert(!tube || D == 3, "the tube predicate is 3-D");
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int Q = D * NT / 2;  // 16-byte pieces of one segment's coefficients
  constexpr int U = kTmU;
  extern __shared__ __align__(16) unsigned char tm_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool want_acc = (!FEAS) && p.sampling_times != nullptr;
  const TmLayout L = tm_layout(D, NT, want_acc, tube, fast_c);
  unsigned char* wbase = tm_smem + (size_t)warp * L.per_warp;
  double* tau_s = reinterpret_cast<double*>(wbase);
  double* stage = reinterpret_cast<double*>(wbase + L.off_stage);
  int4* info_s = reinterpret_cast<int4*>(wbase + L.off_info);
  double* acc_s = reinterpret_cast<double*>(wbase + L.off_acc);
  unsigned char* slots = wbase + L.off_slots;

  const int first = (blockIdx.x * (blockDim.x >> 5) + warp) * 32;  // local index of lane 0's trajectory
  if (first >= p.nb) return;
  const int local = first + lane;
  const bool valid = local < p.nb;
  const int b = p.b0 + (valid ? local : p.nb - 1);
  const size_t B = (size_t)p.B;
  const int K = p.K;
  const size_t S = (size_t)p.max_samples;
  const size_t rec_c = (size_t)K * D * p.N;

  // ---- phase-1 state (lane = trajectory)
  uint32_t st = 0;
  int n = 0, i = 0;
  const double t0 = p.t_start[b], t1 = p.t_end[b], dt = p.dt[b];
  double acc = 0.0, tau = 0.0, Ti = 0.0;
  bool done = !locate_start<AOS_IN>(p, b, t0, dt, i, acc);
  if (done)
    st |= 4u;
  else
    tau = t0 - acc;
  if (!valid) done = true;
  double mv2 = 0.0, ma2 = 0.0;  // FEAS: running maxima of |v|^2, |a|^2 (lane = trajectory)
  unsigned all_bits = 7u;

  // ---- segment slots of this lane's trajectory
  int held0 = -1, held1 = -1;  // segment resident (or in flight) in slot 0 / 1
  int age0 = -1, age1 = -1;    // chunk index whose commit group carries that fetch
  int chunk = 0;
  unsigned char* my_slots = slots + (size_t)lane * L.traj_bytes;
  const double* my_coeffs = p.coeffs + (size_t)b * rec_c;
  const double* my_times = p.seg_times + (size_t)b * K;
  const double* my_geom = tube ? geom + (size_t)(valid ? local : p.nb - 1) * K * kTubeGeomLd : nullptr;
  auto ensure = [&](int seg) {  // make segment `seg` resident in slot seg & 1 (asynchronously)
    const int sl = seg & 1;
    if ((sl ? held1 : held0) == seg) return;
    if (sl) { held1 = seg; age1 = chunk; } else { held0 = seg; age0 = chunk; }
    double* dst = reinterpret_cast<double*>(my_slots + sl * L.slot_bytes);
    const double* src = my_coeffs + (size_t)seg * (D * NT);
#pragma unroll
    for (int q = 0; q < Q; ++q) cp_async16(dst + 2 * q, src + 2 * q);
    if (tube) {
      const double* g = my_geom + (size_t)seg * kTubeGeomLd;
#pragma unroll
      for (int q = 0; q < kTubeGeomLd / 2; ++q) cp_async16(dst + D * NT + 2 * q, g + 2 * q);
    }
    cp_async8(reinterpret_cast<double*>(my_slots + sl * L.slot_bytes + L.t_off), my_times + seg);
  };
  // duration of segment `seg`; with slots: from its landed record (fetching it on demand)
  auto duration = [&](int seg) -> double {
    if (!fast_c) return p.seg_times[at<AOS_IN>((size_t)seg, (size_t)K, B, (size_t)b)];
    const int sl = seg & 1;
    if ((sl ? held1 : held0) != seg) {
      ensure(seg);
      cp_async_commit();
      cp_async_wait_all();
    } else if ((sl ? age1 : age0) >= chunk - 1) {
      cp_async_wait_all();  // carried by the newest group: not retired by wait_group 1
    }
    return *reinterpret_cast<const double*>(my_slots + sl * L.slot_bytes + L.t_off);
  };
  if (fast_c) {
    if (!done) {
      ensure(i);
      if (i + 1 < K) ensure(i + 1);
    }
    cp_async_commit();
    ++chunk;
  }

  for (;; ++chunk) {
    // ------------------------------------------------ phase 1: trajectory.cpp:114-133
    if (fast_c) cp_async_wait_group1();
    int cnt = 0, seg0 = i, cross = kTmChunk;  // samples [cross, cnt) lie in segment seg0 + 1
    if (!done) {
      Ti = duration(i);
      const int limit = min(kTmChunk, p.max_samples - n);
      double* trow = tau_s + lane * kTmTauLd;
      double* arow = acc_s + lane * kTmTauLd;
      for (;;) {
        // straight run inside the current segment
        while (cnt < limit && acc < t1 && !(tau > Ti)) {
          trow[cnt] = tau;
          if (want_acc) arow[cnt] = acc;
          tau += dt;
          acc += dt;
          ++cnt;
        }
        if (!(acc < t1)) {
          done = true;
          break;
        }
        if (tau > Ti) {  // crossing: no sample emitted
          tau = tau - Ti;
          ++i;
          if (i >= K) {
            done = true;
            break;
          }
          if (cnt == 0) {
            seg0 = i;
          } else if (i > seg0 + 1) {
            break;  // a third segment: leave it to the next chunk
          } else {
            cross = cnt;
          }
          Ti = duration(i);
          continue;
        }
        // cnt == limit
        if (limit < kTmChunk) {  // out of output rows
          st |= 8u;
          done = true;
        }
        break;
      }
    }
    info_s[lane] = make_int4(cnt, n, seg0, cross);
    __syncwarp();
    if (!__any_sync(FULL, cnt > 0)) break;

    // ------------------------------------------------ phase 2: lane = sample, U trajectories at a time
#pragma unroll 1
    for (int r = 0; r < 32; r += U) {
      int4 info[U];
#pragma unroll
      for (int u = 0; u < U; ++u) info[u] = info_s[r + u];
      bool any = false;
#pragma unroll
      for (int u = 0; u < U; ++u) any = any || info[u].x > 0;
      if (!any) continue;
      double x[U][D];
      int seg[U];
      bool act[U];
      unsigned fl[U];
      double wv[U], wa[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int cnt_r = info[u].x;
        act[u] = lane < cnt_r;
        const int l = act[u] ? lane : max(cnt_r - 1, 0);
        const double ta = tau_s[(r + u) * kTmTauLd + l];
        seg[u] = info[u].z + (l >= info[u].w ? 1 : 0);
        double c[D][NT];
        const double2* slot =
            reinterpret_cast<const double2*>(slots + (size_t)(r + u) * L.traj_bytes + (seg[u] & 1) * L.slot_bytes);
        if (fast_c) {
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const double2 v = slot[q];
            c[(2 * q) / NT][(2 * q) % NT] = v.x;
            c[(2 * q + 1) / NT][(2 * q + 1) % NT] = v.y;
          }
        } else {
          const int br = min(p.b0 + first + r + u, p.b0 + p.nb - 1);
          const double* src = p.coeffs + at<AOS_IN>((size_t)seg[u] * D * p.N, rec_c, B, (size_t)br);
          const size_t stride = AOS_IN ? 1 : B;
#pragma unroll
          for (int dim = 0; dim < D; ++dim)
#pragma unroll
            for (int j = 0; j < NT; ++j) c[dim][j] = (j < p.N) ? __ldg(src + ((size_t)dim * p.N + j) * stride) : 0.0;
        }
        if (!FEAS) {
          if (MODE == TM_POSITION) {
#pragma unroll
            for (int dim = 0; dim < D; ++dim) x[u][dim] = horner<NT>(c[dim], ta);
          } else {
            // polynomial.h:136-149 with the table row B(derivative, .)
            const int der = p.derivative;
#pragma unroll
            for (int dim = 0; dim < D; ++dim) {
              double acc_h = 0.0;
#pragma unroll
              for (int j = NT - 1; j >= 0; --j)
                if (j >= der) acc_h = fma(acc_h, ta, c_tab.base[der * MTG_BASE_LD + j] * c[dim][j]);
              x[u][dim] = acc_h;
            }
          }
        } else {
          double v2 = 0.0, a2 = 0.0;
#pragma unroll
          for (int dim = 0; dim < D; ++dim) {
            double p0 = c[dim][NT - 1], p1 = 0.0, p2 = 0.0;
#pragma unroll
            for (int j = NT - 2; j >= 0; --j) {
              p2 = fma(p2, ta, p1);
              p1 = fma(p1, ta, p0);
              p0 = fma(p0, ta, c[dim][j]);
            }
            x[u][dim] = p0;
            v2 = fma(p1, p1, v2);
            a2 = fma(2.0 * p2, 2.0 * p2, a2);
          }
          // sqrt is monotone and correctly rounded: |v| <= v_max <=> |v|^2 <= v2_lim (host-computed
          // largest double whose root is <= v_max), and max|v| = sqrt(max |v|^2).
          unsigned f = (v2 <= p.v2_lim ? 1u : 0u) | (a2 <= p.a2_lim ? 2u : 0u) | 4u;
          if (D == 3 && tube) {
            const int lr = min(first + r + u, p.nb - 1);
            const double2* g = fast_c ? slot + Q
                                      : reinterpret_cast<const double2*>(geom + ((size_t)lr * K + seg[u]) * kTubeGeomLd);
            const double2 g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3], g4 = g[4], g5 = g[5], g6 = g[6], g7 = g[7];
            TubeSeg t;
            t.A[0] = g0.x; t.A[1] = g0.y; t.A[2] = g1.x; t.A[3] = g1.y; t.A[4] = g2.x; t.A[5] = g2.y;
            t.bvec[0] = g3.x; t.bvec[1] = g3.y; t.bvec[2] = g4.x;
            t.n[0] = g4.y; t.n[1] = g5.x; t.n[2] = g5.y;
            t.cs = g6.x; t.ce = g6.y; t.r2 = g7.x;
            const double x3[3] = {x[u][0], x[u][D > 1 ? 1 : 0], x[u][D > 2 ? 2 : 0]};
            if (!in_tube(t, x3)) f &= 3u;
          }
          if (!act[u]) {
            f = 7u;
            v2 = 0.0;
            a2 = 0.0;
          }
          fl[u] = f;
          wv[u] = v2;
          wa[u] = a2;
        }
      }
      if (FEAS) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const double mv = warp_max_nonneg(wv[u]), ma = warp_max_nonneg(wa[u]);
          const unsigned wf = __reduce_and_sync(FULL, fl[u]);
          if (lane == r + u) {
            mv2 = fmax(mv2, mv);
            ma2 = fmax(ma2, ma);
            all_bits &= wf;
          }
          if (p.flags && act[u]) p.flags[(size_t)(p.b0 + first + r + u) * S + info[u].y + lane] = (uint8_t)fl[u];
        }
      }
      if (p.samples) {
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int dim = 0; dim < D; ++dim) stage[u * (32 * D) + lane * D + dim] = x[u][dim];
        __syncwarp();
#pragma unroll
        for (int u = 0; u < U; ++u) {
          double* out = p.samples + ((size_t)(p.b0 + first + r + u) * S + info[u].y) * D;
          const int total = info[u].x * D;
#pragma unroll
          for (int q = 0; q < D; ++q) {
            const int e = lane + 32 * q;
            if (e < total) out[e] = stage[u * (32 * D) + e];
          }
        }
        __syncwarp();
      }
      if (!FEAS) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (act[u]) {
            const size_t o = (size_t)(p.b0 + first + r + u) * S + info[u].y + lane;
            if (want_acc) p.sampling_times[o] = acc_s[(r + u) * kTmTauLd + lane];
            if (p.segment_idx) p.segment_idx[o] = seg[u];
          }
      }
    }
    n += cnt;
    __syncwarp();
    // every sample emitted so far has been evaluated: both slots may be re-targeted.
    // Prefetch the current and the next segment (no-ops while they are resident).
    if (fast_c) {
      if (!done) {
        ensure(i);
        if (i + 1 < K) ensure(i + 1);
      }
      cp_async_commit();
    }
  }
  cp_async_wait_all();
  if (valid) {
    if (p.n_samples) p.n_samples[b] = n;
    if (p.status) p.status[b] = st;
    if (FEAS) {
      if (p.max_v) p.max_v[b] = sqrt(mv2);
      if (p.max_a) p.max_a[b] = sqrt(ma2);
      if (p.feasible) p.feasible[b] = (uint8_t)((all_bits == 7u && st == 0) ? 1 : 0);
    }
  }
}

}  // namespace mtg
#endif

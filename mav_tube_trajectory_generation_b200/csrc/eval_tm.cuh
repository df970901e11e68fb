// Sampled evaluation / feasibility sweep with TRAJECTORY-CONTIGUOUS outputs ("tm" kernels): the
// reference's own order — evaluateRange fills one std::vector<VectorXd> per trajectory
// (trajectory.cpp:74-134).
//
// One WARP owns TPW = 16 trajectories and alternates two phases over chunks of 32 samples:
//
//  phase 1  lane = trajectory. Every lane replays the reference's SERIAL sampling recurrence
//           (acc += dt; tau += dt; tau -= T_i on a strict '>' crossing, which emits no sample) for
//           up to 32 samples of its trajectory — a whole 8-sample block per trip while no crossing
//           or end is near, then four, then one (rounding is monotone, so the last sample's two
//           tests cover the others). The position / derivative sweeps park only tau of each block's
//           FIRST sample (the phase-2 lane replays the block's adds: same operations, same bits); the
//           feasibility sweeps, which are register-bound, park every tau (and acc goes to shared
//           memory whenever sampling_times is wanted). This is the part of the algorithm that cannot
//           be parallelised over samples: ~4 fp64 ops per sample.
//  phase 2  4 lanes x 8 consecutive samples per trajectory, 8 trajectories per pass. A lane keeps
//           its segment's coefficients in registers for its 8 samples (re-broadcasting them per
//           sample costs 240 B/sample of shared-memory bandwidth: the wall an earlier lane =
//           sample version hit) and advances the 8 Horner evaluations together: 24 independent FMA
//           chains for D = 3. The blocks of a chunk are laid out so that none straddles a segment
//           crossing (segment A owns the first ceil(cross/8) blocks, segment B starts a new one),
//           so there is no divergence. Results are staged in a bank-skewed shared-memory tile and
//           leave the warp as whole, consecutive 256-byte stores: D*32 doubles of ONE trajectory
//           per row. (Storing the registers directly as 16-byte pieces writes half sectors and
//           measured 36 % slower.)
//
// Segment records {coefficients, tube constants, duration} live in per-trajectory shared-memory
// SLOT PAIRS: segment s sits in slot s & 1, is fetched ONCE with cp.async (LDGSTS) by the
// trajectory's lane, and the next segment is prefetched right after a chunk has been evaluated, a
// few chunks before it is needed, so neither phase waits on DRAM in steady state (one commit group
// per chunk; wait_group 1 retires all but the newest). A chunk stops early if a third segment
// would start (tiny segments or large dt).
//
// The warp is self-contained (only __syncwarp; one warp per CTA), so occupancy is set by the
// shared memory of a warp alone: 15.9 kB for the position sweep (13 warps/SM).
// Replaces (reference): Polynomial::evaluate polynomial.h:136-149, Segment::evaluate
// segment.cpp:51-58, Trajectory::evaluateRange trajectory.cpp:74-134, the sampled limit check
// test_utils.h:43-54 / NL_I:2686-2733 and the sampled form of the tube geometry
// polynomial_optimization_qcqp_impl.h:357-474.
#ifndef MTG_EVAL_TM_CUH_
#define MTG_EVAL_TM_CUH_

#include <stdint.h>

#include "eval.cuh"

namespace mtg {

constexpr int kTmChunk = 32;              // samples per trajectory per chunk (= lanes)
constexpr int kTmTauLd = kTmChunk + 1;    // odd stride in 8-byte words: conflict-free lane-major stores
constexpr int kTubeGeomLd = 16;           // doubles per (trajectory, segment) tube record (one 128-B line)

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
// Every per-sample output is trajectory-contiguous: x[b * max_samples * width + n * width + ...].
//
// Shared memory per warp (TmLayout): tau[TPW][33] | staging tile [8][4*(8 D + 1)] | info, offsets,
// counts | flag rows | acc[TPW][33] (only when sampling_times is requested) | slot pairs [TPW].
#ifndef MTG_TM_STORE_MODE
#define MTG_TM_STORE_MODE 0  // sample rows: 0 plain stores, 1 streaming (st.cs), 2 write-through (st.wt); measured, see profiles/r02_experiments
#endif
#if MTG_TM_STORE_MODE == 1
#define MTG_TM_STORE(ptr, v) __stcs((ptr), (v))
#elif MTG_TM_STORE_MODE == 2
#define MTG_TM_STORE(ptr, v) __stwt((ptr), (v))
#else
#define MTG_TM_STORE(ptr, v) (*(ptr) = (v))
#endif
#ifndef MTG_TM_FEAS_MINB
#define MTG_TM_FEAS_MINB 8  // resident warps per SM the feasibility instantiations are compiled for
#endif
#ifndef MTG_TM_FEAS_JB
#define MTG_TM_FEAS_JB 4    // samples a lane advances together in the feasibility sweep (3 * JB * D FMA chains)
#endif
#ifndef MTG_TM_TAUBLK
#define MTG_TM_TAUBLK 1  // 1: phase 1 parks tau at block starts only, phase-2 lanes replay their 8 adds
#endif
// (position / derivative sweeps only: measured +6 % there, -10 % on the register-bound feasibility sweep)
__host__ __device__ constexpr bool tm_taublk(int mode) { return MTG_TM_TAUBLK != 0 && mode < 2; }
constexpr int kTmBlkLd = 5;              // parked block-start taus per trajectory (4) + 1: odd stride
constexpr int kTmR = 8;                  // consecutive samples per lane in phase 2
constexpr int kTmG = 32 / (kTmChunk / kTmR);  // trajectories per phase-2 pass (8): 4 lanes each
struct TmLayout {
  int slot_bytes;   // one segment record: D*NT coefficients (+ tube record); multiple of 16
  int traj_bytes;   // 2 slots + {T of slot 0, T of slot 1}
  int blk_ld;       // doubles per lane block in the staging tile: 8*D + 1 (bank skew)
  int row_ld;       // doubles per trajectory row of the staging tile: 4 * blk_ld
  int off_stage, off_info, off_off, off_cnt, off_flag, off_acc, off_slots, off_dt, per_warp;
};
// TPW = trajectories per warp (phase-1 lanes in use): the shared memory of a warp scales with it.
// 16 (half the phase-1 lanes idle) doubles the resident warps against 32 and is what every mode ships with; for the
// feasibility sweep 8 and 32 were measured and lose (profiles/r02_experiments, profiles/r01_sweep_ab_experiments.log).
#ifndef MTG_TM_FEAS_TPW
#define MTG_TM_FEAS_TPW 16
#endif
__host__ __device__ constexpr int tm_tpw(int mode) { return mode >= 2 ? MTG_TM_FEAS_TPW : 16; }
__host__ __device__ inline TmLayout tm_layout(int D, int NT, bool want_acc, bool tube, int kTmTPW, int mode) {
  const bool kTmTauBlk = tm_taublk(mode);
  TmLayout L;
  L.slot_bytes = D * NT * 8 + (tube ? kTubeGeomLd * 8 : 0);
  L.traj_bytes = 2 * L.slot_bytes + 16;
  L.blk_ld = kTmR * D + 1;
  L.row_ld = (kTmChunk / kTmR) * L.blk_ld;
  L.off_dt = kTmTPW * (kTmTauBlk ? kTmBlkLd : kTmTauLd) * 8;
  L.off_stage = L.off_dt + (kTmTauBlk ? kTmTPW * 8 : 0);
  L.off_info = L.off_stage + kTmG * L.row_ld * 8;
  L.off_off = L.off_info + kTmTPW * 16;
  L.off_cnt = L.off_off + kTmTPW * 8;
  L.off_flag = L.off_cnt + kTmTPW * 4;
  L.off_acc = L.off_flag + kTmG * 40;
  L.off_slots = L.off_acc + (want_acc ? kTmTPW * kTmTauLd * 8 : 0);
  L.per_warp = L.off_slots + kTmTPW * L.traj_bytes;
  L.per_warp = (L.per_warp + 15) & ~15;
  return L;
}

enum TmMode { TM_POSITION = 0, TM_DERIVATIVE = 1, TM_FEAS = 2, TM_FEAS_TUBE = 3 };

// Requirements (checked by the launcher, which otherwise falls back to the one-thread-per-
// trajectory kernels of eval.cuh): AoS layout, N == NT, coeffs 16-byte aligned.
template <int NT, int D, int MODE>
__global__ void __launch_bounds__(32, (MODE >= 2 ? MTG_TM_FEAS_MINB : MODE == 1 ? 12 : 13)) eval_tm_kernel(const EvalParams p, const double* __restrict__ geom) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr bool FEAS = MODE >= TM_FEAS;
  constexpr bool tube = MODE == TM_FEAS_TUBE;
  constexpr int Q = D * NT / 2;  // 16-byte pieces of one segment's coefficients
  constexpr int R = kTmR, G = kTmG, kTmTPW = tm_tpw(MODE);
  static_assert(!tube || D == 3, "the tube predicate is 3-D");
  extern __shared__ __align__(16) unsigned char tm_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr bool EXTRA = MODE == TM_DERIVATIVE;  // sampling_times / segment_idx outputs exist in this mode only
  const bool want_acc = EXTRA && p.sampling_times != nullptr;
  constexpr bool kTmTauBlk = tm_taublk(MODE);
  const TmLayout L = tm_layout(D, NT, want_acc, tube, kTmTPW, MODE);
  unsigned char* wbase = tm_smem + (size_t)warp * L.per_warp;
  double* tau_s = reinterpret_cast<double*>(wbase);
  double* stage = reinterpret_cast<double*>(wbase + L.off_stage);
  int4* info_s = reinterpret_cast<int4*>(wbase + L.off_info);
  unsigned char* flag_s = wbase + L.off_flag;
  size_t* off_s = reinterpret_cast<size_t*>(wbase + L.off_off);
  int* cnt_s = reinterpret_cast<int*>(wbase + L.off_cnt);
  double* acc_s = reinterpret_cast<double*>(wbase + L.off_acc);
  unsigned char* slots = wbase + L.off_slots;
  double* dt_s = reinterpret_cast<double*>(wbase + L.off_dt);
  constexpr int TAU_LD = kTmTauBlk ? kTmBlkLd : kTmTauLd;

  const int first = (blockIdx.x * (blockDim.x >> 5) + warp) * kTmTPW;  // local index of lane 0's trajectory
  if (first >= p.nb) return;
  const int local = first + lane;
  const bool valid = local < p.nb && lane < kTmTPW;
  const int b = p.b0 + (valid ? local : p.nb - 1);
  const int K = p.K;
  const size_t S = (size_t)p.max_samples;

  const size_t traj_off = (size_t)(p.b0 + local) * S;  // first output row of this lane's trajectory
  const int q8 = lane >> 2, sb = lane & 3;  // phase-2 role: trajectory within the pass, block within the trajectory
  int skew[D];  // staging-tile position of element lane + 32 q of a trajectory row
#pragma unroll
  for (int q = 0; q < D; ++q) skew[q] = (lane + 32 * q) + (lane + 32 * q) / (R * D);

  // ---- phase-1 state (lane = trajectory)
  uint32_t st = 0;
  int n = 0, i = 0;
  const double t0 = p.t_start[b], t1 = p.t_end[b], dt = p.dt[b];
  double acc = 0.0, tau = 0.0, Ti = 0.0;
  bool done = !locate_start<true>(p, b, t0, dt, i, acc);
  if (done)
    st |= 4u;
  else
    tau = t0 - acc;
  if (!valid) done = true;
  if (kTmTauBlk && lane < kTmTPW) dt_s[lane] = dt;
  double mv2 = 0.0, ma2 = 0.0;  // FEAS: running maxima of |v|^2, |a|^2 (lane = trajectory)
  unsigned all_bits = 7u;

  // ---- segment slots of this lane's trajectory
  int held0 = -1, held1 = -1;  // segment resident (or in flight) in slot 0 / 1
  int age0 = -1, age1 = -1;    // chunk index whose commit group carries that fetch
  int chunk = 0;
  unsigned char* my_slots = slots + (size_t)lane * L.traj_bytes;
  const double* my_coeffs = p.coeffs + (size_t)b * ((size_t)K * D * NT);
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
    cp_async8(reinterpret_cast<double*>(my_slots + 2 * L.slot_bytes) + sl, my_times + seg);
  };
  // duration of segment `seg` from its landed record (fetching it on demand)
  auto duration = [&](int seg) -> double {
    const int sl = seg & 1;
    if ((sl ? held1 : held0) != seg) {
      ensure(seg);
      cp_async_commit();
      cp_async_wait_all();
    } else if ((sl ? age1 : age0) >= chunk - 1) {
      cp_async_wait_all();  // carried by the newest group: not retired by wait_group 1
    }
    return reinterpret_cast<const double*>(my_slots + 2 * L.slot_bytes)[sl];
  };
  if (!done) {
    ensure(i);
    if (i + 1 < K) ensure(i + 1);
  }
  cp_async_commit();
  ++chunk;

  for (;; ++chunk) {
    // ------------------------------------------------ phase 1: trajectory.cpp:114-133
    cp_async_wait_group1();
    int cnt = 0, seg0 = i, cross = kTmChunk;  // samples [cross, cnt) lie in segment seg0 + 1
    if (!done) {
      Ti = duration(i);
      const int rows_left = p.max_samples - n;
      int limit = min(kTmChunk, rows_left);
      double* trow = tau_s + lane * TAU_LD;
      double* arow = acc_s + lane * kTmTauLd;
      int mark = 0, blk = 0;  // kTmTauBlk: next block-start sample, blocks parked so far
      for (;;) {
        // straight run inside the current segment: a whole 8-sample block per trip while all eight
        // samples are certain, then four per trip (same adds in the same order: tau_k and acc_k are the
        // reference's values; dt > 0 and rounding is monotone, so the last sample's two tests imply the others)
        while (cnt + R <= limit && (!kTmTauBlk || cnt == mark)) {
          double tk[R], ak[R];
          tk[0] = tau;
          ak[0] = acc;
#pragma unroll
          for (int j = 1; j < R; ++j) {
            tk[j] = tk[j - 1] + dt;
            ak[j] = ak[j - 1] + dt;
          }
          if (!((ak[R - 1] < t1) & !(tk[R - 1] > Ti))) break;
          if (kTmTauBlk) {
            trow[blk++] = tau;
            mark += R;
          } else {
#pragma unroll
            for (int j = 0; j < R; ++j) trow[cnt + j] = tk[j];
          }
          if (want_acc) {
#pragma unroll
            for (int j = 0; j < R; ++j) arow[cnt + j] = ak[j];
          }
          tau = tk[R - 1] + dt;
          acc = ak[R - 1] + dt;
          cnt += R;
        }
        while (cnt + 4 <= limit) {
          const double tau1 = tau + dt, acc1 = acc + dt;
          const double tau2 = tau1 + dt, acc2 = acc1 + dt;
          const double tau3 = tau2 + dt, acc3 = acc2 + dt;
          // dt > 0 and rounding is monotone, so tau <= tau1 <= tau2 <= tau3 and likewise acc: the last
          // sample's two tests imply the other six
          if (!((acc3 < t1) & !(tau3 > Ti))) break;
          if (kTmTauBlk) {
            const int dm = mark - cnt;  // >= 0; blocks are 8 samples apart: at most one starts in this trip
            if (dm < 4) {
              trow[blk++] = dm == 0 ? tau : dm == 1 ? tau1 : dm == 2 ? tau2 : tau3;
              mark += R;
            }
          } else {
            trow[cnt] = tau;
            trow[cnt + 1] = tau1;
            trow[cnt + 2] = tau2;
            trow[cnt + 3] = tau3;
          }
          if (want_acc) {
            arow[cnt] = acc;
            arow[cnt + 1] = acc1;
            arow[cnt + 2] = acc2;
            arow[cnt + 3] = acc3;
          }
          tau = tau3 + dt;
          acc = acc3 + dt;
          cnt += 4;
        }
        while (cnt < limit && acc < t1 && !(tau > Ti)) {
          if (kTmTauBlk) {
            if (cnt == mark) {
              trow[blk++] = tau;
              mark += R;
            }
          } else {
            trow[cnt] = tau;
          }
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
            mark = cnt;  // segment B's blocks start at the crossing
            // phase 2 gives segment A the blocks [0, ceil(cross/8)) and segment B the rest:
            // no 8-sample block straddles the crossing (costs < 8 samples of this chunk)
            limit = min(limit, cross + R * (kTmChunk / R - (cross + R - 1) / R));
          }
          Ti = duration(i);
          continue;
        }
        // cnt == limit
        if (cnt >= rows_left) {  // out of output rows
          st |= 8u;
          done = true;
        }
        break;
      }
    }
    if (lane < kTmTPW) {
      info_s[lane] = make_int4(cnt, n, seg0, cross);
      off_s[lane] = traj_off + (size_t)n;
      cnt_s[lane] = cnt;
    }
    __syncwarp();
    if (!__any_sync(FULL, cnt > 0)) break;

    // ------------------------------------------------ phase 2: 4 lanes x 8 consecutive samples per
    // trajectory, 8 trajectories per pass. The lane's coefficients stay in registers for its 8
    // samples; the one lane whose block contains the crossing reloads in the middle.
#pragma unroll 1
    for (int g = 0; g < kTmTPW / G; ++g) {
      const int r = g * G + q8;
      const int4 info = info_s[r];
      if (!__any_sync(FULL, info.x > 0)) continue;
      const int cnt_r = info.x, last = max(cnt_r - 1, 0);
      // blocks of <= 8 consecutive samples: segment A (seg0) owns the first nA blocks from sample 0,
      // segment B (seg0 + 1) the others from sample `cross`
      const int cross_r = min(info.w, cnt_r);
      const int nA = (cross_r + R - 1) / R;
      const bool isB = sb >= nA;
      const int start = isB ? cross_r + (sb - nA) * R : sb * R;
      const int count = min(R, (isB ? cnt_r : cross_r) - start);  // may be <= 0
      const unsigned char* tslots = slots + (size_t)r * L.traj_bytes;
      double c[D][NT];
      TubeSeg tsg;
      auto load_segment = [&](int seg) {
        const double2* slot = reinterpret_cast<const double2*>(tslots + (seg & 1) * L.slot_bytes);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const double2 v = slot[q];
          c[(2 * q) / NT][(2 * q) % NT] = v.x;
          c[(2 * q + 1) / NT][(2 * q + 1) % NT] = v.y;
        }
        if (tube) {
          const double2* gq = slot + Q;
          const double2 g0 = gq[0], g1 = gq[1], g2 = gq[2], g3 = gq[3], g4 = gq[4], g5 = gq[5], g6 = gq[6], g7 = gq[7];
          tsg.A[0] = g0.x; tsg.A[1] = g0.y; tsg.A[2] = g1.x; tsg.A[3] = g1.y; tsg.A[4] = g2.x; tsg.A[5] = g2.y;
          tsg.bvec[0] = g3.x; tsg.bvec[1] = g3.y; tsg.bvec[2] = g4.x;
          tsg.n[0] = g4.y; tsg.n[1] = g5.x; tsg.n[2] = g5.y;
          tsg.cs = g6.x; tsg.ce = g6.y; tsg.r2 = g7.x;
        }
      };
      load_segment(info.z + (isB ? 1 : 0));
      double v2m = 0.0, a2m = 0.0;
      unsigned fand = 7u;
      double* srow = stage + q8 * L.row_ld;
      // JB samples advance together, one Horner step at a time: JB*D (position) or 3*JB*D
      // (feasibility) independent FMA chains cover the fp64 pipe latency from a single warp.
      constexpr int JB = FEAS ? MTG_TM_FEAS_JB : R;
      double tcur = 0.0, dt_r = 0.0;
      if (kTmTauBlk) {
        tcur = tau_s[r * TAU_LD + sb];
        dt_r = dt_s[r];
      }
      double x[JB][D];
#pragma unroll
      for (int j0 = 0; j0 < R; j0 += JB) {
        double ta[JB];
        if (kTmTauBlk) {
          // the block's taus from its parked first one: the same adds in the same order as phase 1
#pragma unroll
          for (int j = 0; j < JB; ++j) {
            ta[j] = tcur;
            tcur += dt_r;
          }
        } else {
#pragma unroll
          for (int j = 0; j < JB; ++j) ta[j] = tau_s[r * kTmTauLd + min(start + j0 + j, last)];
        }
        if (!FEAS) {
          if (MODE == TM_POSITION) {
#pragma unroll
            for (int j = 0; j < JB; ++j)
#pragma unroll
              for (int dim = 0; dim < D; ++dim) x[j][dim] = c[dim][NT - 1];
#pragma unroll
            for (int jj = NT - 2; jj >= 0; --jj)
#pragma unroll
              for (int j = 0; j < JB; ++j)
#pragma unroll
                for (int dim = 0; dim < D; ++dim) x[j][dim] = fma(x[j][dim], ta[j], c[dim][jj]);
          } else {
            // polynomial.h:136-149 with the table row B(derivative, .)
            const int der = p.derivative;
#pragma unroll
            for (int j = 0; j < JB; ++j)
#pragma unroll
              for (int dim = 0; dim < D; ++dim) x[j][dim] = 0.0;
#pragma unroll
            for (int jj = NT - 1; jj >= 0; --jj) {
              if (jj >= der) {
                double bc[D];
#pragma unroll
                for (int dim = 0; dim < D; ++dim) bc[dim] = c_base.base[der * MTG_BASE_LD + jj] * c[dim][jj];
#pragma unroll
                for (int j = 0; j < JB; ++j)
#pragma unroll
                  for (int dim = 0; dim < D; ++dim) x[j][dim] = fma(x[j][dim], ta[j], bc[dim]);
              }
            }
          }
        } else {
          double p1[JB][D], p2[JB][D];
#pragma unroll
          for (int j = 0; j < JB; ++j)
#pragma unroll
            for (int dim = 0; dim < D; ++dim) {
              x[j][dim] = c[dim][NT - 1];
              p1[j][dim] = 0.0;
              p2[j][dim] = 0.0;
            }
#pragma unroll
          for (int jj = NT - 2; jj >= 0; --jj)
#pragma unroll
            for (int j = 0; j < JB; ++j)
#pragma unroll
              for (int dim = 0; dim < D; ++dim) {
                p2[j][dim] = fma(p2[j][dim], ta[j], p1[j][dim]);
                p1[j][dim] = fma(p1[j][dim], ta[j], x[j][dim]);
                x[j][dim] = fma(x[j][dim], ta[j], c[dim][jj]);
              }
#pragma unroll
          for (int j = 0; j < JB; ++j) {
            double v2 = 0.0, a2 = 0.0;
#pragma unroll
            for (int dim = 0; dim < D; ++dim) {
              v2 = fma(p1[j][dim], p1[j][dim], v2);
              a2 = fma(2.0 * p2[j][dim], 2.0 * p2[j][dim], a2);
            }
            // sqrt is monotone and correctly rounded: |v| <= v_max <=> |v|^2 <= v2_lim (host-computed
            // largest double whose root is <= v_max), and max|v| = sqrt(max |v|^2).
            unsigned f = (v2 <= p.v2_lim ? 1u : 0u) | (a2 <= p.a2_lim ? 2u : 0u) | 4u;
            if (tube) {
              const double x3[3] = {x[j][0], x[j][D > 1 ? 1 : 0], x[j][D > 2 ? 2 : 0]};
              if (!in_tube(tsg, x3)) f &= 3u;
            }
            if (j0 + j < count) {
              v2m = fmax(v2m, v2);
              a2m = fmax(a2m, a2);
              fand &= f;
              flag_s[q8 * 40 + start + j0 + j] = (unsigned char)f;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < JB; ++j) {
          const int k = start + j0 + j;
          if (j0 + j < count) {
#pragma unroll
            for (int dim = 0; dim < D; ++dim) srow[k * D + (k >> 3) + dim] = x[j][dim];
          }
        }
      }
      if (FEAS) {
        // block maxima -> trajectory maxima (4 lanes) -> the trajectory's phase-1 lane
#pragma unroll
        for (int m = 1; m <= 2; m <<= 1) {
          v2m = fmax(v2m, __shfl_xor_sync(FULL, v2m, m));
          a2m = fmax(a2m, __shfl_xor_sync(FULL, a2m, m));
          fand &= __shfl_xor_sync(FULL, fand, m);
        }
        const int src = 4 * (lane & (G - 1));
        const double ov = __shfl_sync(FULL, v2m, src), oa = __shfl_sync(FULL, a2m, src);
        const unsigned of = __shfl_sync(FULL, fand, src);
        if ((lane / G) == g) {
          mv2 = fmax(mv2, ov);
          ma2 = fmax(ma2, oa);
          all_bits &= of;
        }
      }
      __syncwarp();
      // staged rows -> global memory: whole consecutive 256-byte stores per trajectory
      // (cnt == 0 rows fall out through the predicates; everything else is branch-free)
#pragma unroll
      for (int t = 0; t < G; ++t) {
        const int cnt_t = cnt_s[g * G + t];
        const size_t o = off_s[g * G + t];
        if (p.samples) {
          double* out = p.samples + o * D + lane;
          const double* row = stage + t * L.row_ld;
          const int total = cnt_t * D;
#pragma unroll
          for (int q = 0; q < D; ++q)
            if (lane + 32 * q < total) MTG_TM_STORE(out + 32 * q, row[skew[q]]);
        }
        if (FEAS) {
          if (p.flags && lane < cnt_t) p.flags[o + lane] = flag_s[t * 40 + lane];
        } else if (EXTRA) {
          if (lane < cnt_t) {
            const int4 it = info_s[g * G + t];
            if (want_acc) p.sampling_times[o + lane] = acc_s[(g * G + t) * kTmTauLd + lane];
            if (p.segment_idx) p.segment_idx[o + lane] = it.z + (lane >= it.w ? 1 : 0);
          }
        }
      }
      __syncwarp();
    }
    n += cnt;
    // every sample emitted so far has been evaluated: both slots may be re-targeted.
    // Prefetch the current and the next segment (no-ops while they are resident).
    if (!done) {
      ensure(i);
      if (i + 1 < K) ensure(i + 1);
    }
    cp_async_commit();
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

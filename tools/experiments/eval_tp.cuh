// NOT BUILT, NOT SHIPPED: the software-pipelined sweep kernel of round 1, kept as the record of a measured-and-
// rejected design. It passed the full parity suite (tests/test_eval_gpu.py, test_sweep_gpu.py)
// bit for bit and ran at 0.97x of csrc/eval_tm.cuh (profiles/r01_sweep_ab_experiments.log, ab_pipe2). To build it
// again: copy next to csrc/eval_tm.cuh and launch eval_tp_kernel with tp_layout(...).per_warp bytes per warp.
// Sampled evaluation / feasibility sweep, trajectory-contiguous outputs, SOFTWARE-PIPELINED ("tp"):
// the successor of eval_tm.cuh (same outputs bit for bit, same parameters).
//
// eval_tm alternates two phases per chunk of 32 samples: the reference's serial sampling recurrence
// (trajectory.cpp:114-133; lane = trajectory, 16 of 32 lanes, one dependent fp64 add per sample,
// divergent at segment crossings) and the Horner evaluation (4 lanes x 8 samples per trajectory).
// ncu: the recurrence is 41 % of the instructions but 63 % of the warp's time (latency and branch
// bound). Here the recurrence of chunk c + 1 is written BRANCH-FREE, in blocks of <= 8 samples
// ("trips"), and placed in the same basic blocks as the Horner evaluation of chunk c, so the
// scheduler interleaves its dependent add / compare chain with the 24 independent FMA chains:
//
//   trip (lane = trajectory): end test, at most one segment crossing (tau -= T_i; ++i), then up to 8
//     predicated steps {tau += dt; acc += dt} while (acc < t_end) & !(tau > T_i) & rows are left.
//     It parks tau of the block's first sample and the block's sample count; a block never
//     straddles a crossing, because the crossing is taken at the START of the next trip.
//     4 trips = 1 chunk of <= 32 samples in <= 2 segments (a second crossing ends the chunk).
//   pass (4 lanes x 1 block per trajectory, 8 trajectories): the lane replays its block's adds from
//     the parked tau (same operations in the same order: bit-identical), keeps its segment's
//     coefficients in registers for the 8 samples, stages the results in a bank-skewed tile and the
//     warp stores whole 256-byte rows (as in eval_tm).
//
// Segment coefficient records live in a ring of 3 shared-memory slots per trajectory (slot = seg % 3),
// fetched with cp.async one iteration before the chunk that needs them is evaluated; segment
// durations travel in registers (T_i, T_{i+1}) and are loaded straight from global memory.
// One warp per CTA, only __syncwarp; occupancy is set by the shared memory of a warp (~20 kB).
// Replaces (reference): Polynomial::evaluate polynomial.h:136-149, Segment::evaluate
// segment.cpp:51-58, Trajectory::evaluateRange trajectory.cpp:74-134, the sampled limit check
// test_utils.h:43-54 / NL_I:2686-2733 and the sampled tube geometry QC_I:357-474.
#ifndef MTG_EVAL_TP_CUH_
#define MTG_EVAL_TP_CUH_

#include <stdint.h>

#include "eval_tm.cuh"

namespace mtg {

constexpr int kTpTPW = 16;   // trajectories per warp
constexpr int kTpRing = 3;   // segment slots per trajectory
struct TpLayout {
  int slot_bytes, traj_bytes, blk_ld, row_ld;
  int off_tb, off_dt, off_info, off_stage, off_flag, off_acc, off_slots, per_warp;
};
__host__ __device__ inline TpLayout tp_layout(int D, int NT, bool want_acc, bool tube, bool feas) {
  const bool half = !feas;  // position / derivative sweeps stage 4 trajectories at a time
  TpLayout L;
  L.slot_bytes = D * NT * 8 + (tube ? kTubeGeomLd * 8 : 0);
  L.traj_bytes = kTpRing * L.slot_bytes;  // 720 / 1104 B = 20 words (mod 32): 8 trajectories, 8 distinct 16-byte bank groups
  L.blk_ld = kTmR * D + 1;
  L.row_ld = (kTmChunk / kTmR) * L.blk_ld;
  L.off_tb = 0;                                              // [2][TPW][5] block-start taus
  L.off_dt = L.off_tb + 2 * kTpTPW * kTmBlkLd * 8;           // [TPW]
  L.off_info = L.off_dt + kTpTPW * 8;                        // [2][TPW] int4
  L.off_stage = L.off_info + 2 * kTpTPW * 16;                // [G][row_ld]
  L.off_flag = L.off_stage + (half ? kTmG / 2 : kTmG) * L.row_ld * 8;  // [G][40] bytes
  L.off_acc = L.off_flag + (feas ? kTmG * 40 : 0);           // [2][TPW][33] (sampling_times only)
  L.off_acc = (L.off_acc + 15) & ~15;
  L.off_slots = L.off_acc + (want_acc ? 2 * kTpTPW * kTmTauLd * 8 : 0);
  L.per_warp = L.off_slots + kTpTPW * L.traj_bytes;
  L.per_warp = (L.per_warp + 15) & ~15;
  return L;
}

__device__ __forceinline__ double ldg_if(const double* ptr, bool pred, double otherwise) {
  double v = otherwise;
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q ld.global.nc.f64 %0, [%1]; }"
               : "+d"(v)
               : "l"(__cvta_generic_to_global(ptr)), "r"((int)pred));
  return v;
}

template <int NT, int D, int MODE>
__global__ void __launch_bounds__(32, (MODE >= 2 ? 8 : MODE == 1 ? 11 : 12)) eval_tp_kernel(const EvalParams p, const double* __restrict__ geom) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr bool FEAS = MODE >= TM_FEAS;
  constexpr bool tube = MODE == TM_FEAS_TUBE;
  constexpr bool EXTRA = MODE == TM_DERIVATIVE;  // sampling_times / segment_idx outputs exist in this mode only
  constexpr int Q = D * NT / 2;  // 16-byte pieces of one segment's coefficients
  constexpr int R = kTmR, G = kTmG, TPW = kTpTPW;
  static_assert(!tube || D == 3, "the tube predicate is 3-D");
  static_assert(TPW == 2 * G, "two passes per chunk");
  extern __shared__ __align__(16) unsigned char tp_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool want_acc = EXTRA && p.sampling_times != nullptr;
  const TpLayout L = tp_layout(D, NT, want_acc, tube, FEAS);
  unsigned char* wbase = tp_smem + (size_t)warp * L.per_warp;
  double* tb_s = reinterpret_cast<double*>(wbase + L.off_tb);
  double* dt_s = reinterpret_cast<double*>(wbase + L.off_dt);
  int4* info_s = reinterpret_cast<int4*>(wbase + L.off_info);
  double* stage = reinterpret_cast<double*>(wbase + L.off_stage);
  unsigned char* flag_s = wbase + L.off_flag;
  double* acc_s = reinterpret_cast<double*>(wbase + L.off_acc);
  unsigned char* slots = wbase + L.off_slots;

  const int first = (blockIdx.x * (blockDim.x >> 5) + warp) * TPW;  // local index of lane 0's trajectory
  if (first >= p.nb) return;
  const int local = first + lane;
  const bool valid = local < p.nb && lane < TPW;
  const int b = p.b0 + (valid ? local : p.nb - 1);
  const int K = p.K;
  const size_t S = (size_t)p.max_samples;
  const int q8 = lane >> 2, sb = lane & 3;  // pass role: trajectory within the pass, block within the trajectory
  int skew[D];  // staging-tile position of element lane + 32 q of a trajectory row
#pragma unroll
  for (int q = 0; q < D; ++q) skew[q] = (lane + 32 * q) + (lane + 32 * q) / (R * D);

  // ---- recurrence state (lane = trajectory)
  uint32_t st = 0;
  int n = 0, i = 0;
  const double t0 = p.t_start[b], t1 = p.t_end[b], dt = p.dt[b];
  double acc = 0.0, tau = 0.0;
  bool done = !locate_start<true>(p, b, t0, dt, i, acc);
  if (done)
    st |= 4u;
  else
    tau = t0 - acc;
  if (!valid) done = true;
  const double* my_times = p.seg_times + (size_t)b * K;
  double Ti = done ? 0.0 : my_times[i];
  double Tn = (!done && i + 1 < K) ? my_times[i + 1] : 0.0;
  double Tnn = 0.0;  // T_{i_cs + 2}, requested a whole chunk before it can be needed
  if (lane < TPW) dt_s[lane] = dt;
  double mv2 = 0.0, ma2 = 0.0;  // FEAS: running maxima of |v|^2, |a|^2 (lane = trajectory)
  unsigned all_bits = 7u;

  // ---- segment slots of this lane's trajectory: ring of 3, slot = seg % 3
  int held0 = -1, held1 = -1, held2 = -1;
  unsigned char* my_slots = slots + (size_t)lane * L.traj_bytes;
  const double* my_coeffs = p.coeffs + (size_t)b * ((size_t)K * D * NT);
  const double* my_geom = tube ? geom + (size_t)(valid ? local : p.nb - 1) * K * kTubeGeomLd : nullptr;
  // makes segment `seg` resident (asynchronously) unless its slot still holds a segment of the chunk
  // that is waiting to be evaluated (segments pend, pend + 1); true if a fetch was issued
  auto ensure = [&](int seg, int pend) -> bool {
    if (seg >= K) return false;
    const int sl = seg % kTpRing;
    const int h = sl == 0 ? held0 : sl == 1 ? held1 : held2;
    if (h == seg) return false;
    if (h >= 0 && (h == pend || h == pend + 1)) return false;
    if (sl == 0) held0 = seg; else if (sl == 1) held1 = seg; else held2 = seg;
    double* dst = reinterpret_cast<double*>(my_slots + sl * L.slot_bytes);
    const double* src = my_coeffs + (size_t)seg * (D * NT);
#pragma unroll
    for (int q = 0; q < Q; ++q) cp_async16(dst + 2 * q, src + 2 * q);
    if (tube) {
      const double* gsrc = my_geom + (size_t)seg * kTubeGeomLd;
#pragma unroll
      for (int q = 0; q < kTubeGeomLd / 2; ++q) cp_async16(dst + D * NT + 2 * q, gsrc + 2 * q);
    }
    return true;
  };

  // ---- one chunk of the recurrence = chunk_begin, 4 trips, chunk_end (lane = trajectory)
  int counts = 0, rel = 0, i_cs = i, cnt = 0;
  bool stop = false;
  auto chunk_begin = [&]() {
    counts = 0;
    rel = 0;
    cnt = 0;
    Tn = i != i_cs ? Tnn : Tn;  // the last chunk crossed (at most once) and moved Tn into Ti
    i_cs = i;
    Tnn = ldg_if(my_times + min(i + 2, K - 1), !done && i + 2 < K, 0.0);
    stop = false;
  };
  // trajectory.cpp:114-133, <= 8 samples, no branches
  int t_room = 0, t_m = 0;
  double* t_arow = acc_s;
  auto trip_head = [&](int blk, int buf) {
    bool act = !done && !stop;
    const bool ended = act && !(acc < t1);  // the end of the range
    done |= ended;
    act = act && !ended;
    const bool crossing = act && tau > Ti;  // crossing: no sample emitted
    const bool last = i + 1 >= K;
    const bool defer = crossing && !last && i + 1 - i_cs > 1;  // a third segment: leave it to the next chunk
    const bool go = crossing && !defer;
    stop |= defer;
    tau = go ? tau - Ti : tau;
    i += go ? 1 : 0;
    Ti = go ? Tn : Ti;  // (Tn is stale until the next chunk_begin: a second crossing ends the chunk)
    done |= go && last;
    act = act && !defer && !(go && last);
    const int rows_left = p.max_samples - n - cnt;
    const bool norow = act && !(tau > Ti) && rows_left <= 0;  // a sample is due and there is no row for it
    st |= norow ? 8u : 0u;
    done |= norow;
    act = act && !norow;
    if (lane < TPW) tb_s[(buf * TPW + lane) * kTmBlkLd + blk] = tau;
    t_room = act ? min(R, rows_left) : 0;
    t_arow = acc_s + (buf * TPW + lane) * kTmTauLd + cnt;
    t_m = 0;
  };
  // step j: emitted iff every earlier step of the block was (t_m == j), a row is left, acc < t_end and
  // !(tau > T_i); then tau += dt, acc += dt as predicated adds (no selects, no branches)
  auto trip_step = [&](int j) {
    const double acc_prev = acc;
    asm("{\n\t.reg .pred q;\n\t"
        "setp.eq.s32 q, %2, %6;\n\t"
        "setp.lt.and.s32 q, %6, %7, q;\n\t"
        "setp.lt.and.f64 q, %1, %4, q;\n\t"
        "setp.leu.and.f64 q, %0, %5, q;\n\t"
        "@q add.rn.f64 %0, %0, %3;\n\t"
        "@q add.rn.f64 %1, %1, %3;\n\t"
        "@q add.s32 %2, %2, 1;\n\t}"
        : "+d"(tau), "+d"(acc), "+r"(t_m)
        : "d"(dt), "d"(t1), "d"(Ti), "r"(j), "r"(t_room));
    if (want_acc && t_m == j + 1) t_arow[j] = acc_prev;
  };
  auto trip_tail = [&](int blk) {
    counts |= t_m << (8 * blk);
    rel |= (i - i_cs) << (2 * blk);
    cnt += t_m;
  };
  auto trip = [&](int blk, int buf) {
    trip_head(blk, buf);
#pragma unroll
    for (int j = 0; j < R; ++j) trip_step(j);
    trip_tail(blk);
  };
  auto chunk_end = [&](int buf) {
    if (lane < TPW) info_s[buf * TPW + lane] = make_int4(counts, n, i_cs, rel | (cnt << 8));
    n += cnt;
  };

  // ---- prologue: segments of chunk 0, recurrence of chunk 0, segments of chunk 1
  if (!done) {
    ensure(i, -2);
    ensure(i + 1, -2);
  }
  cp_async_commit();
  chunk_begin();
  int pend = i_cs;  // first segment of the chunk waiting to be evaluated
#pragma unroll
  for (int t = 0; t < 4; ++t) trip(t, 0);
  chunk_end(0);
  bool urgent = false;
  if (!done) {
    ensure(i, pend);
    ensure(i + 1, pend);
  }
  cp_async_commit();

  for (int c = 0;; ++c) {
    const int buf = c & 1;
    if (__any_sync(FULL, urgent))
      cp_async_wait_all();
    else
      cp_async_wait_group1();
    __syncwarp();
    // chunk c: any samples?  (lane = trajectory: cnt of chunk c is still in this lane's info)
    const int my_cnt = lane < TPW ? (info_s[buf * TPW + lane].w >> 8) : 0;
    if (!__any_sync(FULL, my_cnt > 0 || !done)) break;
    const int next_pend = i;  // == i_cs of chunk c + 1
    chunk_begin();

#pragma unroll
    for (int g = 0; g < TPW / G; ++g) {
      // pass g of chunk c, with two blocks of the recurrence of chunk c + 1 threaded through its
      // Horner levels in source order as 20 micro-steps (head, 8 steps, tail per block)
      constexpr int kMicro = 2 * (R + 2);
      auto micro = [&](int idx) {
        const int t = idx / (R + 2), u = idx % (R + 2);
        if (u == 0)
          trip_head(2 * g + t, buf ^ 1);
        else if (u <= R)
          trip_step(u - 1);
        else
          trip_tail(2 * g + t);
      };
      int mdone = 0;  // micro-steps issued so far (a compile-time constant after unrolling)
      auto micro_upto = [&](int upto) {
#pragma unroll
        for (int q = 0; q < kMicro; ++q)
          if (q >= mdone && q < upto) micro(q);
        mdone = max(mdone, min(upto, kMicro));
      };

      const int r = g * G + q8;
      const int4 info = info_s[buf * TPW + r];
      const int c0 = info.x & 255, c1 = (info.x >> 8) & 255, c2 = (info.x >> 16) & 255;
      const int count = (info.x >> (8 * sb)) & 255;
      const int start = (sb > 0 ? c0 : 0) + (sb > 1 ? c1 : 0) + (sb > 2 ? c2 : 0);
      const int seg = info.z + ((info.w >> (2 * sb)) & 3);
      const unsigned char* tslots = slots + (size_t)r * L.traj_bytes;
      double cf[D][NT];
      TubeSeg tsg;
      {
        const double2* slot = reinterpret_cast<const double2*>(tslots + (seg % kTpRing) * L.slot_bytes);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const double2 v = slot[q];
          cf[(2 * q) / NT][(2 * q) % NT] = v.x;
          cf[(2 * q + 1) / NT][(2 * q + 1) % NT] = v.y;
        }
        if (tube) {
          const double2* gq = slot + Q;
          const double2 g0 = gq[0], g1 = gq[1], g2 = gq[2], g3 = gq[3], g4 = gq[4], g5 = gq[5], g6 = gq[6], g7 = gq[7];
          tsg.A[0] = g0.x; tsg.A[1] = g0.y; tsg.A[2] = g1.x; tsg.A[3] = g1.y; tsg.A[4] = g2.x; tsg.A[5] = g2.y;
          tsg.bvec[0] = g3.x; tsg.bvec[1] = g3.y; tsg.bvec[2] = g4.x;
          tsg.n[0] = g4.y; tsg.n[1] = g5.x; tsg.n[2] = g5.y;
          tsg.cs = g6.x; tsg.ce = g6.y; tsg.r2 = g7.x;
        }
      }
      double v2m = 0.0, a2m = 0.0;
      unsigned fand = 7u;
      constexpr bool HALF = !FEAS;
      double* srow = stage + (HALF ? (q8 & (G / 2 - 1)) : q8) * L.row_ld;
      // JB samples advance together, one Horner step at a time: JB*D (position) or 3*JB*D
      // (feasibility) independent FMA chains
      constexpr int JB = FEAS ? 4 : R;
      double tcur = tb_s[(buf * TPW + r) * kTmBlkLd + sb];
      const double dt_r = dt_s[r];
      constexpr int kLevels = (R / JB) * (MODE == TM_DERIVATIVE ? NT : NT - 1);  // Horner levels of this pass
      constexpr int kPer = kMicro / kLevels > 0 ? kMicro / kLevels : 1;          // micro-steps per level
      int lvl = 0;
      double x[JB][D];
#pragma unroll
      for (int j0 = 0; j0 < R; j0 += JB) {
        double ta[JB];
        // the block's taus from its parked first one: the same adds in the same order as the trip
#pragma unroll
        for (int j = 0; j < JB; ++j) {
          ta[j] = tcur;
          tcur += dt_r;
        }
        if (!FEAS) {
          if (MODE == TM_POSITION) {
#pragma unroll
            for (int j = 0; j < JB; ++j)
#pragma unroll
              for (int dim = 0; dim < D; ++dim) x[j][dim] = cf[dim][NT - 1];
#pragma unroll
            for (int jj = NT - 2; jj >= 0; --jj) {
#pragma unroll
              for (int j = 0; j < JB; ++j)
#pragma unroll
                for (int dim = 0; dim < D; ++dim) x[j][dim] = fma(x[j][dim], ta[j], cf[dim][jj]);
              micro_upto(++lvl * kPer);
            }
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
                for (int dim = 0; dim < D; ++dim) bc[dim] = c_tab.base[der * MTG_BASE_LD + jj] * cf[dim][jj];
#pragma unroll
                for (int j = 0; j < JB; ++j)
#pragma unroll
                  for (int dim = 0; dim < D; ++dim) x[j][dim] = fma(x[j][dim], ta[j], bc[dim]);
              }
              micro_upto(++lvl * kPer);
            }
          }
        } else {
          double p1[JB][D], p2[JB][D];
#pragma unroll
          for (int j = 0; j < JB; ++j)
#pragma unroll
            for (int dim = 0; dim < D; ++dim) {
              x[j][dim] = cf[dim][NT - 1];
              p1[j][dim] = 0.0;
              p2[j][dim] = 0.0;
            }
#pragma unroll
          for (int jj = NT - 2; jj >= 0; --jj) {
#pragma unroll
            for (int j = 0; j < JB; ++j)
#pragma unroll
              for (int dim = 0; dim < D; ++dim) {
                p2[j][dim] = fma(p2[j][dim], ta[j], p1[j][dim]);
                p1[j][dim] = fma(p1[j][dim], ta[j], x[j][dim]);
                x[j][dim] = fma(x[j][dim], ta[j], cf[dim][jj]);
              }
            micro_upto(++lvl * kPer);
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
        if (!HALF) {
#pragma unroll
          for (int j = 0; j < JB; ++j) {
            const int k = start + j0 + j;
            if (j0 + j < count) {
#pragma unroll
              for (int dim = 0; dim < D; ++dim) srow[k * D + (k >> 3) + dim] = x[j][dim];
            }
          }
        }
      }
      micro_upto(kMicro);
      if (FEAS) {
        // block maxima -> trajectory maxima (4 lanes) -> the trajectory's recurrence lane
#pragma unroll
        for (int mm = 1; mm <= 2; mm <<= 1) {
          v2m = fmax(v2m, __shfl_xor_sync(FULL, v2m, mm));
          a2m = fmax(a2m, __shfl_xor_sync(FULL, a2m, mm));
          fand &= __shfl_xor_sync(FULL, fand, mm);
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
      if (!HALF) __syncwarp();
      // staged rows -> global memory: whole consecutive 256-byte stores per trajectory
      // (empty rows fall out through the predicates; everything is branch-free)
      int4 its[G];
#pragma unroll
      for (int t = 0; t < G; ++t) its[t] = info_s[buf * TPW + g * G + t];
#pragma unroll
      for (int hh = 0; hh < (HALF ? 2 : 1); ++hh) {
      if (HALF) {
        if ((q8 >> 2) == hh && p.samples) {
#pragma unroll
          for (int j = 0; j < JB; ++j) {
            const int k = start + j;
            if (j < count) {
#pragma unroll
              for (int dim = 0; dim < D; ++dim) srow[k * D + (k >> 3) + dim] = x[j][dim];
            }
          }
        }
        __syncwarp();
      }
      constexpr int GH = HALF ? G / 2 : G;
      if (p.samples) {
        // all loads of the tile first, then the stores: the store of a row does not wait for its own load
        double v[GH][D];
#pragma unroll
        for (int t = 0; t < GH; ++t)
#pragma unroll
          for (int q = 0; q < D; ++q) v[t][q] = stage[t * L.row_ld + skew[q]];
#pragma unroll
        for (int tt = 0; tt < GH; ++tt) {
          const int t = hh * GH + tt;
          const size_t o = (size_t)(p.b0 + first + g * G + t) * S + (size_t)its[t].y;
          double* out = p.samples + o * D + lane;
          const int total = (its[t].w >> 8) * D;
#pragma unroll
          for (int q = 0; q < D; ++q)
            if (lane + 32 * q < total) out[32 * q] = v[tt][q];
        }
      }
      if (HALF) __syncwarp();
      }
      if (FEAS || EXTRA)
#pragma unroll
      for (int t = 0; t < G; ++t) {
        const int4 it = its[t];
        const int cnt_t = it.w >> 8;
        const size_t o = (size_t)(p.b0 + first + g * G + t) * S + (size_t)it.y;
        if (FEAS) {
          if (p.flags && lane < cnt_t) p.flags[o + lane] = flag_s[t * 40 + lane];
        } else if (EXTRA) {
          if (lane < cnt_t) {
            if (want_acc) p.sampling_times[o + lane] = acc_s[(buf * TPW + g * G + t) * kTmTauLd + lane];
            if (p.segment_idx) {
              const int e0 = it.x & 255, e1 = e0 + ((it.x >> 8) & 255), e2 = e1 + ((it.x >> 16) & 255);
              const int blk = (lane >= e0 ? 1 : 0) + (lane >= e1 ? 1 : 0) + (lane >= e2 ? 1 : 0);
              p.segment_idx[o + lane] = it.z + ((it.w >> (2 * blk)) & 3);
            }
          }
        }
      }
      __syncwarp();
    }
    chunk_end(buf ^ 1);
    // chunk c has been evaluated: its slots may be re-targeted. First whatever chunk c + 1 still
    // misses (a fetch that had to wait for chunk c: needed right away), then the segments of chunk c + 2.
    urgent = false;
    if (valid) {
      urgent = ensure(next_pend, -2);
      urgent = ensure(next_pend + 1, -2) || urgent;
    }
    if (!done) {
      ensure(i, next_pend);
      ensure(i + 1, next_pend);
    }
    cp_async_commit();
    pend = next_pend;
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

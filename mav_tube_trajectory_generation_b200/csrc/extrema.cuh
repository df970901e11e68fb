// E5 (analytic) / E6 / R1 — extrema of the derivative magnitude |p^(d)(t)| per segment and per
// trajectory, the candidate lists behind them, and the real roots of a polynomial in an interval.
//
// Replaces (reference): Segment::computeMinMaxMagnitudeCandidateTimes segment.cpp:82-133
// (g = sum_dim conv(delta, delta'), polynomial.cpp:163-181; one dimension: roots of p^(d+1)),
// Polynomial::computeMinMaxCandidates / selectMinMaxCandidatesFromRoots polynomial.cpp:32-83,
// findRootsJenkinsTraub + rpoly_ak1 rpoly_ak1.cpp:57-937, the candidate evaluation
// segment.cpp:135-184 and Trajectory::computeMinMaxMagnitude trajectory.cpp:184-220.
//
// Root finding. The reference runs Jenkins-Traub for ALL complex roots and keeps the real ones inside
// [t_start, t_end] (polynomial.cpp:46-60). Only those are ever consumed, so this kernel isolates exactly them,
// in the BERNSTEIN basis of the interval: the number of sign changes V of the Bernstein coefficients bounds the
// number of roots in the open interval (variation diminishing) — V = 0: none, V = 1: exactly one — and halving
// the interval (de Casteljau) makes the control polygon converge to the curve, so an interval with V >= 2 is
// split until every piece has V <= 1. The single root of a V = 1 piece is then polished by a bracketed Newton
// iteration on the power form, started at the zero crossing of the control polygon (quadratically close to the
// root) and stopped when |g(t)| falls inside its own rounding-error bound. On min-snap segments one split per
// problem is typical (0.8-1.0 on average) against 15 levels x 2-5 brackets for a derivative-chain isolator:
// ~1e3 instead of ~4.5e3 multiply-adds per problem and no level loop. Coefficients below 1e-12 of the largest
// Bernstein coefficient count as zero (1e-14 for an explicit input polynomial): that removes the numerically
// 7-fold root a rest-to-rest segment has AT its end point (the end points are candidates anyway) and can only
// hide a root PAIR whose dip of g is below 1e-12 of its scale — not an extremum of the magnitude at any tolerance
// used here. A root that falls ON a split point makes the shared coefficient negligible for both halves; it is
// recognised by the sign change ACROSS that coefficient and queued directly. An interval that does not start at
// t = 0 is isolated from the origin (two sides) rather than through a Taylor shift, which would lose
// ~(1 + |t_start|)^n in the coefficients. All loops are bounded.
//
// Mapping (warp-cooperative). One warp owns kExG = 16 root problems ((trajectory, segment) pairs, contiguous in
// memory in both layouts); g, the roots and the candidate values of a problem live in shared memory (odd
// strides), nothing in local memory. The derivative coefficients are staged with batched loads (8 in flight per
// lane), g is built by the symmetric form of the convolution (p p' = (p^2)'/2: half the multiply-adds) in
// registers so that the staging area can overlap it, and converted to the Bernstein basis by n rounds of
// "add the left neighbour" over shuffles. Intervals wait on a per-warp stack in shared memory; a group of 16
// lanes (32 for more than 16 coefficients) takes one interval: lane i holds Bernstein coefficient i, V and the
// control-polygon crossing come from ballots, a split is n rounds of shuffles (the left child is the first lane's
// value after every round, the right child is what the lanes hold at the end). V = 1 pieces queue as brackets and
// are polished one per lane. The candidates [t_start, t_end, roots...] are then evaluated one per lane.
// 10.7 kB of shared memory per warp: 20 warps per SM.
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
constexpr int kRootIters = 64;      // Newton iterations per root (bisection fallback inside)
constexpr int kBernDepth = 30;      // halvings of one interval before it is taken as it is
constexpr int kExG = 16;            // root problems per warp
#ifndef MTG_EX_WARPS
#define MTG_EX_WARPS 10
#endif
constexpr int kExWarps = MTG_EX_WARPS;  // warps per CTA
constexpr int kExBr = 32;           // bracket queue of a warp
#ifndef MTG_EX_MINB
#define MTG_EX_MINB 2                // resident CTAs the register allocation aims for (shared memory allows 2)
#endif
static_assert(kExG == 16, "extrema_warp_kernel maps problem = lane & 15");

struct ExtremaParams {
  const double* __restrict__ coeffs;     // elem ((i*D + dim)*N + j), rec K*D*N; raw mode: elem j, rec N
  const double* __restrict__ seg_times;  // elem i, rec K (nullptr in raw mode with explicit bounds)
  const double* __restrict__ t_lo;       // elem i, rec K; or nullptr = 0
  const double* __restrict__ t_hi;       // elem i, rec K; or nullptr = the segment time
  double* __restrict__ seg_out;          // scratch [nb][K][4]: min_t, min_v, max_t, max_v (chunk-local); or nullptr
  uint32_t* __restrict__ seg_status;     // scratch [nb][K]; or nullptr
  double* __restrict__ min_value;        // [B] or nullptr
  double* __restrict__ min_time;         // [B] or nullptr (relative to the segment start, extremum.h:41-42)
  int32_t* __restrict__ min_seg;         // [B] or nullptr
  double* __restrict__ max_value;
  double* __restrict__ max_time;
  int32_t* __restrict__ max_seg;
  double* __restrict__ seg_max_value;    // elem i, rec K; or nullptr: per-segment maxima (candidates of LIN_I:455-487)
  double* __restrict__ seg_max_time;     // elem i, rec K; or nullptr
  double* __restrict__ cand_time;        // elem i*max_cand + q, rec K*max_cand; or nullptr
  double* __restrict__ cand_value;       // same; or nullptr
  int32_t* __restrict__ n_cand;          // elem i, rec K; or nullptr
  uint32_t* __restrict__ status;         // [B] or nullptr
  int B, b0, nb, K, N, D, derivative;
  int max_cand;
  int dim_mask;  // bit per dimension that takes part in the magnitude (segment.cpp:82-86 `dimensions`)
  int raw;       // 1: coeffs ARE the polynomial whose roots are wanted (findRootsJenkinsTraub), cand_* = the roots
  // soft constraint (NL_I:2735-2766): soft_cost[b] (+)= min(soft_max, exp((max - soft_limit) / soft_limit * soft_weight))
  double* __restrict__ soft_cost;       // [B] or nullptr
  double* __restrict__ soft_violation;  // [B] or nullptr: max - soft_limit (evaluateMaximumMagnitudeConstraint, NL_I:2686-2733)
  double soft_limit, soft_weight, soft_max;
  int soft_accumulate;
};

// shared-memory plan of one launch (host and device agree through these numbers)
struct ExtremaPlan {
  int len;     // coefficients of g
  int S;       // stride (doubles) of the per-problem arrays g, roots, values
  int nd;      // coefficients of p^(d)
  int sd;      // stride (doubles) of the staged derivative coefficients of one problem
  int ndim;    // dimensions taking part
  int lpi;     // lanes per interval (16 or 32)
  int qc;      // interval stack capacity
  int tld;     // leading dimension of the halving-weight table
  size_t warp_bytes, cta_bytes;
};

__host__ __device__ constexpr ExtremaPlan extrema_plan(int N, int D, int derivative, int dim_mask, int raw) {
  ExtremaPlan pl{};
  int ndim = 0;
  for (int q = 0; q < D; ++q) ndim += (dim_mask >> q) & 1;
  pl.ndim = ndim;
  pl.nd = raw ? N : N - derivative;
  pl.len = raw ? N : (ndim > 1 ? 2 * pl.nd - 2 : pl.nd - 1);
  if (pl.len < 1) pl.len = 1;
  // every per-problem array holds <= len + 1 doubles (len coefficients; len - 1 roots + 2 end points);
  // [g | roots] together must also hold the staged derivative coefficients (D * nd) while g is built
  int S = (pl.len + 1) | 1;
  pl.sd = (raw ? N : D * pl.nd) | 1;  // odd: the 16 problems of a warp then read 16 different banks
  const int need = ((pl.sd + 1) / 2) | 1;
  if (need > S) S = need;
  pl.S = S;
  pl.lpi = pl.len <= 16 ? 16 : 32;
  // the interval stack also stages the scaled coefficients while the first intervals are
  // pushed (at its top end: slot k must not reach the coefficients of problems > k, see the kernel) and the
  // derivative coefficients for the candidate evaluation at the end
  int qc = 32;
  const int stage = raw ? 0 : (kExG * pl.sd + pl.lpi - 1) / pl.lpi;
  if (stage > qc) qc = stage;
  const int scaled = (kExG * S + (kExG - 1) * (pl.lpi > S ? pl.lpi - S : 0) + pl.lpi - 1) / pl.lpi;
  if (scaled > qc) qc = scaled;
  pl.qc = qc;
  // doubles: g, roots [G][S]; lo, hi, eps [G]; stack coefficients [qc][lpi], a, b [qc]; brackets a, b, t [kExBr]
  // ints: n, nroot, st, trajectory, segment [G]; stack meta [qc]; bracket meta [kExBr]; top, nbr
  size_t bytes = (size_t)(2 * kExG * S + 3 * kExG + qc * pl.lpi + 2 * qc + 3 * kExBr) * sizeof(double) +
                 (size_t)(5 * kExG + qc + kExBr + 2) * sizeof(int);
  pl.warp_bytes = (bytes + 15) & ~(size_t)15;
  // CTA-wide tables: B(k, j) and the halving weights C(i, j) / 2^i (odd leading dimension)
  pl.tld = pl.lpi == 16 ? 17 : 23;
  pl.cta_bytes = pl.warp_bytes * kExWarps + (size_t)MTG_BASE_LD * MTG_BASE_LD * sizeof(double) +
                 (size_t)(pl.tld - 1) * pl.tld * sizeof(double);
  return pl;
}

}  // namespace mtg
struct mtg_ctx;
namespace mtg {
// extrema.cu: chunked launch of extrema_warp_kernel (+ extrema_reduce_kernel when per-trajectory outputs are wanted)
int launch_extrema(mtg_ctx* ctx, bool aos, const ExtremaParams& p, cudaStream_t s);

// CN > 0: the instantiation for (N, D, derivative) = (CN, CD, CDER), all dimensions, segment extrema (not raw):
// every size of the shared-memory plan is a compile-time constant then and the small loops unroll.
template <bool AOS, int CN = 0, int CD = 0, int CDER = 0>
__global__ void __launch_bounds__(kExWarps * 32, MTG_EX_MINB) extrema_warp_kernel(const ExtremaParams p_in) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr bool FIX = CN > 0;
  constexpr ExtremaPlan plc = extrema_plan(FIX ? CN : 2, FIX ? CD : 1, FIX ? CDER : 0, FIX ? (1 << CD) - 1 : 1, 0);
  ExtremaParams p = p_in;
  if (FIX) {  // what the launcher has checked; lets the compiler fold every use below
    p.N = CN;
    p.D = CD;
    p.derivative = CDER;
    p.dim_mask = (1 << CD) - 1;
    p.raw = 0;
  }
  const ExtremaPlan pl = FIX ? plc : extrema_plan(p.N, p.D, p.derivative, p.dim_mask, p.raw);
  const int len = pl.len, S = pl.S, nd = pl.nd, G = kExG, LPI = pl.lpi, QC = pl.qc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr unsigned FULL = 0xffffffffu;

  // ---- CTA-wide: the base table B(k, j) = j!/(j-k)! in shared memory (lanes read different rows)
  double* s_base = reinterpret_cast<double*>(smem_raw + pl.warp_bytes * kExWarps);
  for (int i = threadIdx.x; i < MTG_BASE_LD * MTG_BASE_LD; i += blockDim.x) s_base[i] = c_base.base[i];
  // halving weights P[i][j] = C(i, j) / 2^i (exact): de Casteljau at the midpoint in closed form,
  // left child l_i = sum_{j <= i} P[i][j] c_j, right child r_i = sum_{j >= i} P[n-i][j-i] c_j
  double* s_pas = s_base + MTG_BASE_LD * MTG_BASE_LD;
  const int TLD = pl.tld;
  for (int e = threadIdx.x; e < (TLD - 1) * TLD; e += blockDim.x) {
    const int i = e / TLD, j = e - i * TLD;
    double w = 0.0;
    if (j <= i) w = scalbn(c_base.base[j * MTG_BASE_LD + i] / c_base.base[j * MTG_BASE_LD + j], -i);
    s_pas[e] = w;
  }
  __syncthreads();

  // ---- this warp's problems
  const int K = p.K, D = p.D, N = p.N, d = p.derivative;
  const long long n_prob = (long long)p.nb * K;
  const int gps = (p.nb + G - 1) / G;  // SoA: groups per segment
  const long long n_groups = AOS ? (n_prob + G - 1) / G : (long long)gps * K;
  const long long group = (long long)blockIdx.x * kExWarps + warp;
  if (group >= n_groups) return;  // whole warp; no CTA-wide barrier below
  // problem q of the group -> (local trajectory, segment)
  int np;
  int seg0 = 0, local0 = 0;
  long long flat0 = 0;
  if (AOS) {
    flat0 = group * G;
    np = (int)min((long long)G, n_prob - flat0);
  } else {
    seg0 = (int)(group / gps);
    local0 = (int)(group % gps) * G;
    np = min(G, p.nb - local0);
  }
  auto prob_local = [&](int q) -> int { return AOS ? (int)((flat0 + q) / K) : local0 + q; };
  auto prob_seg = [&](int q) -> int { return AOS ? (int)((flat0 + q) % K) : seg0; };

  unsigned char* base_ptr = smem_raw + pl.warp_bytes * warp;
  double* s_g = reinterpret_cast<double*>(base_ptr);  // power coefficients of g; at the end the candidate values
  double* s_root = s_g + G * S;                       // roots found, unsorted until the end
  double* s_lo = s_root + G * S;
  double* s_hi = s_lo + G;
  double* s_eps = s_hi + G;                           // coefficients below this count as zero
  double* s_qc = s_eps + G;                           // interval stack: Bernstein coefficients [QC][LPI]
  double* s_qa = s_qc + QC * LPI;                     //   interval ends
  double* s_qb = s_qa + QC;
  double* s_ba = s_qb + QC;                           // bracket queue: ends and first iterate
  double* s_bb = s_ba + kExBr;
  double* s_bt = s_bb + kExBr;
  int* s_n = reinterpret_cast<int*>(s_bt + kExBr);    // degree of g
  int* s_nroot = s_n + G;
  int* s_st = s_nroot + G;
  int* s_pb = s_st + G;                               // trajectory (batch index) and segment of every problem
  int* s_ps = s_pb + G;
  int* s_qm = s_ps + G;                               // stack meta: problem | depth << 8
  int* s_bm = s_qm + QC;                              // bracket meta: problem | (g < 0 left of the root) << 8
  int* s_top = s_bm + kExBr;
  int* s_nbr = s_top + 1;

  const size_t Bsz = (size_t)p.B;
  const size_t rec_c = p.raw ? (size_t)N : (size_t)K * D * N;
  const int rec_one = p.raw ? N : D * N;  // coefficients of one problem

  // ---- interval of every problem
  if (lane < G) {
    double lo = 0.0, hi = 1.0;
    int b = 0, seg = 0;
    if (lane < np) {
      b = p.b0 + prob_local(lane);
      seg = prob_seg(lane);
      const size_t o = at<AOS>((size_t)seg, (size_t)K, Bsz, (size_t)b);
      lo = p.t_lo ? p.t_lo[o] : 0.0;
      hi = p.t_hi ? p.t_hi[o] : p.seg_times[o];
    }
    s_pb[lane] = b;
    s_ps[lane] = seg;
    s_lo[lane] = lo;
    s_hi[lane] = hi;
    s_st[lane] = 0;
    s_nroot[lane] = 0;
  }
  if (lane == 0) {
    *s_top = 0;
    *s_nbr = 0;
  }

  // Stages the derivative coefficients delta[dim][j] = B(d, j+d) c[j+d] (polynomial.h:99-113) of all problems at
  // `dst` (stride sd per problem, dims not in dim_mask zeroed). Element e = lane + 32 k walks memory in order in both
  // layouts — (problem, coefficient) and the source address advance incrementally — and the loads of kStageU
  // elements are issued before the first is used (one memory round trip per batch).
  // In raw mode the record is the polynomial itself.
  constexpr int kStageU = 8;
  auto stage_delta = [&](double* dst, int sd) {
    const int n_el = np * rec_one;
    if (p.raw) {
      const float inv_a = 1.0f / (float)(AOS ? rec_one : np);
      for (int e = lane; e < n_el; e += 32) {
        int q, r;
        if (AOS) {
          q = (int)(((float)e + 0.5f) * inv_a);
          r = e - q * rec_one;
        } else {
          r = (int)(((float)e + 0.5f) * inv_a);
          q = e - r * np;
        }
        dst[q * sd + r] = p.coeffs[at<AOS>((size_t)r, rec_c, Bsz, (size_t)s_pb[q])];
      }
      return;
    }
    // AoS: the 16 problems are one contiguous run of memory ((b K + seg) is linear in the problem index);
    // SoA: coefficient r of problem q sits at (seg0 rec_one + r) B + b
    int q, r, dq, dr;
    const double* src;
    if (AOS) {
      q = lane / rec_one;
      r = lane - q * rec_one;
      dq = 32 / rec_one;
      dr = 32 - dq * rec_one;
      src = p.coeffs + ((size_t)p.b0 * K + (size_t)flat0) * rec_one + lane;
    } else {
      r = lane / np;
      q = lane - r * np;
      dr = 32 / np;
      dq = 32 - dr * np;
      src = p.coeffs + ((size_t)seg0 * rec_one + r) * Bsz + (size_t)(p.b0 + local0 + q);
    }
    const size_t src_step = AOS ? 32 : (size_t)dr * Bsz + dq;
    const size_t src_wrap = AOS ? 0 : Bsz - np;  // SoA: q wraps to the next coefficient row
    for (int e0 = lane; e0 < n_el; e0 += 32 * kStageU) {
      double v[kStageU];
      int di[kStageU], rv[kStageU];
#pragma unroll
      for (int u = 0; u < kStageU; ++u) {
        v[u] = 0.0;
        di[u] = -1;
        rv[u] = 0;
        if (e0 + 32 * u < n_el) {
          int dim = 0, jj = r;
          while (jj >= N) {
            jj -= N;
            ++dim;
          }
          if (jj >= d) {
            v[u] = *src;
            di[u] = q * sd + dim * nd + jj - d;
            rv[u] = dim | (jj << 8);
          }
        }
        src += src_step;
        if (AOS) {
          q += dq;
          r += dr;
          if (r >= rec_one) {
            r -= rec_one;
            ++q;
          }
        } else {
          r += dr;
          q += dq;
          if (q >= np) {
            q -= np;
            ++r;
            src += src_wrap;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kStageU; ++u) {
        if (di[u] < 0) continue;
        const int dim = rv[u] & 255, jj = rv[u] >> 8;
        dst[di[u]] = ((p.dim_mask >> dim) & 1) ? s_base[d * MTG_BASE_LD + jj] * v[u] : 0.0;
      }
    }
  };

  // ---- g: the polynomial whose real roots in [lo, hi] are the candidate times
  {
    double* s_delta = s_g;  // [g | roots] holds D * nd doubles per problem until g replaces them
    const int sd = pl.sd;
    stage_delta(s_delta, sd);
    __syncwarp();
    // lane = (problem q, coefficient m = 2 k + half): 16 problems x 2 lanes, conflict-free odd strides; the
    // coefficients wait in registers until every lane has finished reading delta
    const int q = lane & 15;
    double acc[kMaxG / 2];
    if (q < np) {
      const double* dl = s_delta + q * sd;
#pragma unroll
      for (int k = 0; k < kMaxG / 2; ++k) {
        const int m = 2 * k + (lane >> 4);
        double a = 0.0;
        if (m < len) {
          if (p.raw) {
            a = dl[m];
          } else if (pl.ndim > 1) {
            // sum_dim conv(delta, delta'), delta'[j] = (j+1) delta[j+1] (segment.cpp:93-115), summed by symmetry:
            // p p' = (p^2)' / 2, so g[m] = (m+1) (sum_{i < j, i + j = m+1} delta_i delta_j + delta_{(m+1)/2}^2 / 2)
            const int k2 = m + 1;
            const int i0 = max(0, k2 - (nd - 1)), i1 = (k2 - 1) >> 1;
            for (int dim = 0; dim < D; ++dim) {
              const double* x = dl + dim * nd;
              for (int i = i0; i <= i1; ++i) a = fma(x[i], x[k2 - i], a);
              if (!(k2 & 1)) a = fma(0.5 * x[k2 >> 1], x[k2 >> 1], a);
            }
            a *= (double)k2;
          } else {
            // one dimension: roots of p^(d+1)   (segment.cpp:124-131); the other dimensions are staged as zeros
            for (int dim = 0; dim < D; ++dim) a += (double)(m + 1) * dl[dim * nd + m + 1];
          }
        }
        acc[k] = a;
      }
    }
    __syncwarp();
    if (q < np) {
#pragma unroll
      for (int k = 0; k < kMaxG / 2; ++k) {
        const int m = 2 * k + (lane >> 4);
        if (m < len) s_g[q * S + m] = acc[k];
      }
    }
    __syncwarp();
  }
  // strip zero leading coefficients (findLastNonZeroCoeff, rpoly_ak1.cpp:57-68); an interval that is empty or not
  // finite has no roots either
  {
    int n = -1;
    if (lane < np) {
      n = len - 1;
      while (n >= 0 && !(fabs(s_g[lane * S + n]) >= 2.2250738585072014e-308)) --n;
      const double L = s_hi[lane] - s_lo[lane];
      if (!(L > 0.0) || !(L < 1.7e308)) n = min(n, 0);
    }
    if (lane < G) s_n[lane] = n;
    __syncwarp();
  }
  const int GP = 32 / LPI;            // intervals per warp step
  const int gi = lane / LPI, li = lane - gi * LPI;

  // Newton polish of the queued brackets, one per lane: value, slope and the running rounding-error bound of the
  // value from one coefficient stream (p and p' by the coupled Horner recurrence, err = sum |c_j| |t|^j)
  auto drain = [&]() {
    const int nbr = *s_nbr;
    for (int e = lane; e < nbr; e += 32) {
      const int meta = s_bm[e];
      const int q = meta & 255;
      const bool fa_neg = (meta >> 8) & 1;
      const int n = s_n[q];
      const double* gq = s_g + q * S;
      double a = s_ba[e], bb = s_bb[e], t = s_bt[e];
      const double tol = 1e-15 * (s_hi[q] - s_lo[q]);
      const double wmin = fmax(tol, 4.5e-16 * fmax(fabs(s_lo[q]), fabs(s_hi[q])));
      double res = t;
      for (int it = 0;; ++it) {
        double ft = gq[n], dft = 0.0, err = fabs(ft);
        const double at = fabs(t);
#pragma unroll 4
        for (int j = n - 1; j >= 0; --j) {
          const double c = gq[j];
          dft = fma(dft, t, ft);
          ft = fma(ft, t, c);
          err = fma(err, at, fabs(c));
        }
        res = t;
        // |value| inside its own rounding noise: the root is located as well as fp64 can tell
        if (fabs(ft) <= (double)(2 * n + 2) * 1.1102230246251565e-16 * err) break;
        if ((ft < 0.0) == fa_neg)
          a = t;
        else
          bb = t;
        if (!(bb - a > wmin)) break;
        // Newton step with a cheap reciprocal (rcp.approx.f64 + one Newton step: ~1e-12 relative — the step only
        // has to land inside the bracket; anything else, incl. NaN / infinity from a vanishing slope, bisects)
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(dft));
        r = r * fma(-dft, r, 2.0);
        double tn = fma(-ft, r, t);
        if (!(tn > a && tn < bb)) tn = 0.5 * (a + bb);
        if (!(fabs(tn - t) > tol)) {
          res = tn;
          break;
        }
        t = tn;
        res = t;
        if (it + 1 >= kRootIters) {
          atomicOr(&s_st[q], 16);  // MTG_ST_NO_CONVERGENCE (reference: rpoly returns partial roots, RPOLY_C:372-377)
          break;
        }
      }
      const int slot = atomicAdd(&s_nroot[q], 1);
      if (slot < S) s_root[q * S + slot] = res;
    }
    __syncwarp();
    if (lane == 0) *s_nbr = 0;
    __syncwarp();
  };

  // ---- isolate. The Bernstein coefficients of g on an interval that starts at t = 0 come straight from the power
  // coefficients: b_i = sum_{j <= i} [C(i,j) / C(n,j)] g_j E^j (C(i,j)/C(n,j) = B(j,i)/B(j,n)), E the far end — as
  // well conditioned as evaluating g. A Taylor shift to t_start is NOT (it loses ~(1 + |t_start|)^n), so an
  // interval [lo, hi] is covered from the origin instead: side 0 isolates on [0, hi], side 1 on [lo, 0] (E = lo,
  // coefficients stored in reverse so that the interval still runs left to right); pieces outside [lo, hi] are
  // dropped when they are popped and roots outside it at the end. Segment extrema (lo = 0) only have side 0.
  if constexpr (FIX) {
    // ---- lane = interval (compile-time sizes; t_start = 0). A lane holds the Bernstein coefficients of ITS interval
    // in registers: counting V, the control-polygon crossing and the halving are serial loops over <= NC
    // coefficients, 32 intervals at a time. A split keeps the right half in the lane and pushes the left half to the
    // warp's stack, where the next idle lane picks it up (in the first round lanes 16-31 are idle: they take the
    // first left halves at once).
    constexpr int NC = plc.len, NM = NC - 1, SL = NC | 1;
    const int QS = min(QC, (QC * LPI) / SL);  // stack slots of SL doubles (the meta arrays hold QC)
    double* s_sc = s_qc;             // scaled power coefficients [problem][S], before the stack is used
    {
      // every problem is taken at the full degree NM (a stripped leading zero is a zero power coefficient: the
      // Bernstein form of the degree-elevated polynomial, for which the variation count holds just the same)
      const int q = lane & 15;
      if (q < np && s_n[q] >= 1) {
        const double E = s_hi[q];
        double lp = (lane >> 4) ? E : 1.0;
        const double E2 = E * E;
#pragma unroll
        for (int j2 = 0; j2 < NC; j2 += 2) {
          const int j = j2 + (lane >> 4);
          if (j < NC) s_sc[q * S + j] = s_g[q * S + j] * lp * (s_base[j * MTG_BASE_LD + j] / s_base[j * MTG_BASE_LD + NM]);  // 1 / C(NM, j)
          lp *= E2;
        }
      }
    }
    __syncwarp();
    double c[NC];
    bool busy = false;
    int q = lane & 15, depth = 0;
    constexpr int n = NM;
    double a = 0.0, b = 0.0, eps = 0.0;
    if (lane < np && s_n[lane] >= 1) {
      busy = true;
#pragma unroll
      for (int j = 0; j < NC; ++j) c[j] = s_sc[lane * S + j];
      // b_i = sum_{j <= i} C(i,j) s_j: the s_j are the forward differences of the b_i at 0
#pragma unroll
      for (int r = 1; r <= NM; ++r)
#pragma unroll
        for (int i = NM; i >= r; --i)
          c[i] += c[i - 1];
      double mx = 0.0;
#pragma unroll
      for (int j = 0; j < NC; ++j)
        mx = fmax(mx, fabs(c[j]));
      eps = 1e-12 * mx;  // "zero" for the sign-variation count (see the header)
      s_eps[lane] = eps;
      b = s_hi[lane];
      if (c[0] == 0.0) s_root[lane * S + atomicAdd(&s_nroot[lane], 1)] = 0.0;  // a root exactly on an end
      if (c[NM] == 0.0) s_root[lane * S + atomicAdd(&s_nroot[lane], 1)] = b;
    }
    __syncwarp();  // every row has been read: the area is the stack from here on
    for (;;) {
      // ---- idle lanes take intervals off the stack
      const unsigned idle = __ballot_sync(FULL, !busy);
      const int top = *s_top;
      if (idle == FULL && top == 0) break;
      {
        const int rank = __popc(idle & ((1u << lane) - 1u));
        if (!busy && rank < top) {
          const int sl = top - 1 - rank;
          const int meta = s_qm[sl];
          q = meta & 255;
          depth = meta >> 8;
          eps = s_eps[q];
          a = s_qa[sl];
          b = s_qb[sl];
#pragma unroll
          for (int j = 0; j < NC; ++j) c[j] = s_qc[sl * SL + j];
          busy = true;
        }
      }
      __syncwarp();
      const int top1 = top - min(__popc(idle), top);
      // ---- sign variations of the non-negligible coefficients; the first crossing of the control polygon
      int V = 0, i0 = 0, j0 = 1, pidx = 0;
      double ci = 1.0, cj = -1.0, prev = 0.0;
      bool seen = false, first_neg = false;
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const double x = c[j];
        if (busy && fabs(x) > eps) {
          if (seen && ((x < 0.0) != (prev < 0.0))) {
            if (V == 0) {
              i0 = pidx;
              j0 = j;
              ci = prev;
              cj = x;
            }
            ++V;
          }
          if (!seen) first_neg = x < 0.0;
          seen = true;
          prev = x;
          pidx = j;
        }
      }
      // ---- V >= 2: halve, if the left half still has a slot on the stack
      const unsigned want = __ballot_sync(FULL, busy && V >= 2 && depth < kBernDepth);
      const int srank = __popc(want & ((1u << lane) - 1u));
      const bool split = ((want >> lane) & 1u) && top1 + srank < QS;
      const bool leaf = busy && V >= 1 && !split;
      if (busy && V >= 2 && !split && depth < kBernDepth) atomicOr(&s_st[q], 16);  // no room: taken as it is
      // ---- V = 1 (or given up): queue the bracket
      {
        const unsigned lm = __ballot_sync(FULL, leaf);
        if (*s_nbr + __popc(lm) > kExBr) drain();
        if (leaf) {
          double rc;
          const double dc = ci - cj;
          asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(dc));
          const double u = ((double)i0 + ci * rc * (double)(j0 - i0)) * (double)__frcp_rn((float)n);
          double t = a + u * (b - a);
          if (!(t > a && t < b)) t = 0.5 * (a + b);
          const int e = *s_nbr + __popc(lm & ((1u << lane) - 1u));
          s_ba[e] = a;
          s_bb[e] = b;
          s_bt[e] = t;
          s_bm[e] = q | ((first_neg ? 1 : 0) << 8);  // sign of g just right of a
        }
        __syncwarp();
        if (lane == 0) *s_nbr += __popc(lm);
        __syncwarp();
      }
      bool midroot = false, mid_neg = false;
      double mid = 0.0;
      if (split) {
        // de Casteljau at the midpoint without the halvings: after round r, c[i] = 2^r b_i^(r); the left half is
        // c[0] of every round, the right half what is left at the end (index i finished in round n - i)
        const int sl = top1 + srank;
        mid = 0.5 * (a + b);
        s_qa[sl] = a;
        s_qb[sl] = mid;
        s_qm[sl] = q | ((depth + 1) << 8);
        double* ls = s_qc + sl * SL;
        ls[0] = c[0];
        bool lseen = fabs(c[0]) > eps, lneg = c[0] < 0.0;  // the last non-negligible coefficient of the left half
        double shared = 0.0;                                 // b_0^(n) = g(mid): the coefficient both halves share
        double sc = 0.5;
#pragma unroll
        for (int r = 1; r <= NM; ++r) {
#pragma unroll
          for (int i = 0; i + r <= NM; ++i) c[i] += c[i + 1];
          const double lv = c[0] * sc;
          ls[r] = lv;
          if (r < NM) {
            if (fabs(lv) > eps) {
              lseen = true;
              lneg = lv < 0.0;
            }
          } else {
            shared = lv;
          }
          sc *= 0.5;
        }
        // right half: c[i] *= 2^-(NM - i)
        double pw = 1.0 / (double)(1 << NM);
        bool rseen = false, rneg = false;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          c[i] *= pw;
          if (i >= 1 && !rseen && fabs(c[i]) > eps) {
            rseen = true;
            rneg = c[i] < 0.0;
          }
          pw *= 2.0;
        }
        // a root ON the split point: the shared coefficient is negligible for both halves and neither would see the
        // sign change that runs through it — queued directly, bracketed by one control-point spacing on either side
        midroot = !(fabs(shared) > eps) && lseen && rseen && (lneg != rneg);
        mid_neg = lneg;
      }
      {
        const unsigned mm = __ballot_sync(FULL, midroot);
        if (mm) {
          if (*s_nbr + __popc(mm) > kExBr) drain();
          if (midroot) {
            const double h = (b - a) / (double)(2 * n);
            const int e = *s_nbr + __popc(mm & ((1u << lane) - 1u));
            s_ba[e] = mid - h;
            s_bb[e] = mid + h;
            s_bt[e] = mid;
            s_bm[e] = q | ((mid_neg ? 1 : 0) << 8);
          }
          __syncwarp();
          if (lane == 0) *s_nbr += __popc(mm);
        }
      }
      if (split) {  // go on with the right half
        a = mid;
        ++depth;
      } else {
        busy = false;
      }
      __syncwarp();
      if (lane == 0) *s_top = top1 + min(__popc(want), max(QS - top1, 0));
      __syncwarp();
    }
  } else {
    const int nsides = (p.t_lo != nullptr) ? 2 : 1;
    // scaled power coefficients, at the top end of the (empty) stack: pass q0 pushes slots <= q0 + GP - 1 after it has
    // read its own problems' coefficients, and extrema_plan() sizes the stack so that slot k ends below problem k + 1
    double* s_sc = s_qc + QC * LPI - G * S;
    for (int side = 0; side < nsides; ++side) {
      {
        const int q = lane & 15;
        const int n = q < np ? s_n[q] : -1;
        const double E = side == 0 ? s_hi[q] : s_lo[q];
        if (n >= 1 && (side == 0 ? E > 0.0 : E < 0.0)) {
          double lp = (lane >> 4) ? E : 1.0;
          const double E2 = E * E;
          for (int j = lane >> 4; j <= n; j += 2) {
            s_sc[q * S + j] = s_g[q * S + j] * lp * (s_base[j * MTG_BASE_LD + j] / s_base[j * MTG_BASE_LD + n]);  // 1 / C(n, j)
            lp *= E2;
          }
        }
      }
      __syncwarp();
      for (int q0 = 0; q0 < np; q0 += GP) {
        const int q = q0 + gi;
        int n = q < np ? s_n[q] : -1;
        if (n >= 1 && !(side == 0 ? s_hi[q] > 0.0 : s_lo[q] < 0.0)) n = -1;
        // b_i = sum_{j <= i} C(i,j) s_j: the s_j are the forward differences of the b_i at 0, and n rounds of
        // "add the left neighbour" (round r: lanes i >= r) rebuild the values from them
        double c = (n >= 1 && li <= n) ? s_sc[q * S + li] : 0.0;
        {
          int nmax = n;
  #pragma unroll
          for (int m = 16; m >= 1; m >>= 1) nmax = max(nmax, __shfl_xor_sync(FULL, nmax, m));
          for (int r = 1; r <= nmax; ++r) {
            const double dn = __shfl_up_sync(FULL, c, 1, LPI);
            if (li >= r && li <= n) c += dn;
          }
        }
        // scale of the problem: the largest coefficient
        double mx = fabs(c);
  #pragma unroll
        for (int m = 8; m >= 1; m >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, m));
        if (LPI == 32) mx = fmax(mx, __shfl_xor_sync(FULL, mx, 16));
        int slot = 0;
        if (n >= 1 && li == 0) slot = atomicAdd(s_top, 1);  // one slot per interval, drawn by the group's first lane
        const int sl = __shfl_sync(FULL, slot, gi * LPI);
        if (n >= 1) {
          if (li <= n) s_qc[sl * LPI + (side == 0 ? li : n - li)] = c;
          if (li == 0) {
            s_qa[sl] = side == 0 ? 0.0 : s_lo[q];
            s_qb[sl] = side == 0 ? s_hi[q] : 0.0;
            s_qm[sl] = q;
            // "zero" for the sign-variation count: solved trajectories carry ~1e-13 of coefficient noise (the
            // multiple root at a rest-to-rest end is not exact in the data), exact input polynomials only rounding
            s_eps[q] = (p.raw ? 1e-14 : 1e-12) * mx;
            // a root exactly at t = 0 (once: side 1 leaves it to side 0 when both run)
            if (c == 0.0 && (side == 0 || !(s_hi[q] > 0.0))) s_root[q * S + atomicAdd(&s_nroot[q], 1)] = 0.0;
          }
          if (li == n && c == 0.0)  // ... exactly on the far end
            s_root[q * S + atomicAdd(&s_nroot[q], 1)] = side == 0 ? s_hi[q] : s_lo[q];
        }
        __syncwarp();
      }

      // pop intervals, count sign variations, split or queue
      for (;;) {
        __syncwarp();
        const int top = *s_top;
        if (top == 0) break;
        const int take = min(GP, top);
        __syncwarp();
        if (lane == 0) *s_top = top - take;
        const bool have = gi < take;
        const int sl = top - 1 - gi;
        double c = 0.0, a = 0.0, b = 0.0, eps = 0.0;
        int q = 0, depth = 0, n = 0;
        if (have) {
          const int meta = s_qm[sl];
          q = meta & 255;
          depth = meta >> 8;
          n = s_n[q];
          a = s_qa[sl];
          b = s_qb[sl];
          eps = s_eps[q];
          if (li <= n) c = s_qc[sl * LPI + li];
        }
        const bool live = have && !(b < s_lo[q] || a > s_hi[q]);  // a piece outside [lo, hi] is dropped
        __syncwarp();  // the popped slots may be overwritten by the pushes below
        const unsigned shiftg = LPI == 32 ? 0u : 16u * gi;
        const unsigned lmask = LPI == 32 ? FULL : 0xffffu;
        const unsigned P = (__ballot_sync(FULL, live && li <= n && c > eps) >> shiftg) & lmask;
        const unsigned M = (__ballot_sync(FULL, live && li <= n && c < -eps) >> shiftg) & lmask;
        const unsigned nz = P | M;
        // a sign change starts at i: i is non-zero and the next non-zero coefficient above it has the other sign
        int nxt = -1;
        bool var = false;
        if ((nz >> li) & 1u) {
          const unsigned above = li >= 31 ? 0u : (nz & ~((2u << li) - 1u));
          if (above) {
            nxt = __ffs(above) - 1;
            var = ((P >> li) & 1u) != ((P >> nxt) & 1u);
          }
        }
        const unsigned Vm = (__ballot_sync(FULL, var) >> shiftg) & lmask;
        const int V = __popc(Vm);
        const bool room = *s_top + 2 * GP <= QC;   // read before anybody pushes (uniform)
        const bool leaf = have && V >= 1 && (V == 1 || depth >= kBernDepth || !room);
        const bool split = have && V >= 2 && !leaf;
        if (have && V >= 2 && leaf && li == 0 && !room) atomicOr(&s_st[q], 16);
        // the first crossing of the control polygon: between coefficients i0 and j0
        const int i0 = Vm ? __ffs(Vm) - 1 : 0;
        const int j0 = __shfl_sync(FULL, nxt, gi * LPI + i0);
        const double ci = __shfl_sync(FULL, c, gi * LPI + i0);
        const double cj = __shfl_sync(FULL, c, gi * LPI + max(j0, 0));
        if (leaf && li == 0) {
          // a starting point needs no more than a few digits: approximate reciprocals instead of two divisions
          double rc;
          const double dc = ci - cj;
          asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(dc));
          const double u = ((double)i0 + ci * rc * (double)(j0 - i0)) * (double)__frcp_rn((float)n);
          double t = a + u * (b - a);
          if (!(t > a && t < b)) t = 0.5 * (a + b);
          const int e = atomicAdd(s_nbr, 1);
          s_ba[e] = a;
          s_bb[e] = b;
          s_bt[e] = t;
          s_bm[e] = q | (((M >> (__ffs(nz) - 1)) & 1u) << 8);  // sign of g just right of a
        }
        if (__any_sync(FULL, split)) {
          // de Casteljau at the midpoint in closed form: lane i takes l_i and r_i from the parent's coefficients (still
          // in its stack slot: pushes come after the ballots below) — n + 2 multiply-adds per lane instead of n rounds
          double cur = 0.0, left = 0.0;
          if (split && li <= n) {
            const double* cs = s_qc + sl * LPI;
            const double* wl = s_pas + li * TLD;
            const double* wr = s_pas + (n - li) * TLD - li;
            for (int j = 0; j <= n; ++j) {
              const double cj = cs[j];
              if (j <= li) left = fma(wl[j], cj, left);
              if (j >= li) cur = fma(wr[j], cj, cur);
            }
          }
          // A root ON the split point: the shared coefficient g(mid) is then negligible for both children and neither
          // would see the sign change that runs through it. If the nearest non-negligible coefficients on its two sides
          // differ in sign, the root is queued directly, bracketed by one control-point spacing on either side.
          const unsigned PL = (__ballot_sync(FULL, split && li <= n && left > eps) >> shiftg) & lmask;
          const unsigned ML = (__ballot_sync(FULL, split && li <= n && left < -eps) >> shiftg) & lmask;
          const unsigned PR = (__ballot_sync(FULL, split && li <= n && cur > eps) >> shiftg) & lmask;
          const unsigned MR = (__ballot_sync(FULL, split && li <= n && cur < -eps) >> shiftg) & lmask;
          int basei = 0;
          if (split && li == 0) basei = atomicAdd(s_top, 2);
          basei = __shfl_sync(FULL, basei, gi * LPI);
          if (split) {
            const double mid = 0.5 * (a + b);
            s_qc[basei * LPI + li] = li <= n ? cur : 0.0;         // right child first: the left one is popped first
            s_qc[(basei + 1) * LPI + li] = li <= n ? left : 0.0;
            if (li == 0) {
              s_qa[basei] = mid;
              s_qb[basei] = b;
              s_qm[basei] = q | ((depth + 1) << 8);
              s_qa[basei + 1] = a;
              s_qb[basei + 1] = mid;
              s_qm[basei + 1] = q | ((depth + 1) << 8);
              const unsigned nzl = (PL | ML) & ((1u << n) - 1u);   // left child without the shared coefficient (index n)
              const unsigned nzr = (PR | MR) & ~1u;                // right child without it (index 0)
              if (!(((PL | ML) >> n) & 1u) && nzl && nzr) {
                const int il = 31 - __clz(nzl), ir = __ffs(nzr) - 1;
                const bool negl = (ML >> il) & 1u, negr = (MR >> ir) & 1u;
                if (negl != negr) {
                  const double h = (b - a) / (double)(2 * n);
                  const int e = atomicAdd(s_nbr, 1);
                  s_ba[e] = mid - h;
                  s_bb[e] = mid + h;
                  s_bt[e] = mid;
                  s_bm[e] = q | ((negl ? 1 : 0) << 8);
                }
              }
            }
          }
        }
        __syncwarp();
        if (*s_nbr > kExBr - 2 * GP) drain();
      }
    }
  }
  drain();
  // roots inside [lo, hi], ascending; constant polynomial: none (rpoly_ak1.cpp:76-80)
  if (lane < np) {
    double* rq = s_root + lane * S;
    const int nall = min(s_nroot[lane], S - 2);
    const double lo = s_lo[lane], hi = s_hi[lane];
    int nr = 0;
    for (int i = 0; i < nall; ++i) {
      const double x = rq[i];
      if (!(x >= lo && x <= hi)) continue;
      int j = nr - 1;
      while (j >= 0 && rq[j] > x) {
        rq[j + 1] = rq[j];
        --j;
      }
      rq[j + 1] = x;
      ++nr;
    }
    s_nroot[lane] = nr;
  }
  __syncwarp();

  // ---- raw mode: the roots are the result
  if (p.raw) {
    if (lane < np) {
      const int q = lane;
      const int b = s_pb[q], seg = s_ps[q];
      const double* roots = s_root + q * S;
      const int na = s_nroot[q];
      const size_t rec_k = (size_t)K * p.max_cand;
      int w = 0;
      for (int c = 0; c < na && w < p.max_cand; ++c, ++w)
        if (p.cand_time) p.cand_time[at<AOS>((size_t)seg * p.max_cand + w, rec_k, Bsz, (size_t)b)] = roots[c];
      if (p.n_cand) p.n_cand[at<AOS>((size_t)seg, (size_t)K, Bsz, (size_t)b)] = na;
      uint32_t st = (uint32_t)s_st[q];
      if (na > p.max_cand) st |= 8u;  // MTG_ST_TRUNCATED
      if (p.status) p.status[b] = st;
    }
    return;
  }

  // ---- candidates [t_start, t_end, roots...]: |p^(d)(t)| = sqrt(sum_dim evaluate(t, d)^2)  (segment.cpp:135-158)
  double* s_delta = s_qc;  // the interval stack is empty now
  const int sd = pl.sd;
  stage_delta(s_delta, sd);
  __syncwarp();
  {
    // lane = (problem, candidate parity): 16 problems x 2 lanes
    const int q = lane & 15;
    if (q < np) {
      const double* roots = s_root + q * S;
      double* vals = s_g + q * S;
      const double* dl = s_delta + q * sd;
      const int nc = s_nroot[q] + 2;
      for (int c = lane >> 4; c < nc; c += 2) {
        const double t = (c == 0) ? s_lo[q] : (c == 1) ? s_hi[q] : roots[c - 2];
        double m2 = 0.0;
        for (int dim = 0; dim < D; ++dim) {
          double r = 0.0;
          for (int j = nd - 1; j >= 0; --j) r = fma(r, t, dl[dim * nd + j]);
          m2 = fma(r, r, m2);
        }
        vals[c] = sqrt(m2);
      }
    }
  }
  __syncwarp();
  if (lane < np) {
    const int q = lane;
    const int b = s_pb[q], seg = s_ps[q], local = b - p.b0;
    const double* roots = s_root + q * S;
    const double* vals = s_g + q * S;
    const double lo = s_lo[q], hi = s_hi[q];
    const int na = s_nroot[q];
    double mn_v = 1.7976931348623157e308, mx_v = -1.7976931348623157e308, mn_t = 0.0, mx_t = 0.0;
    const size_t rec_k = (size_t)K * p.max_cand;
    int w = 0;
    for (int c = 0; c < na + 2; ++c) {
      const double t = (c == 0) ? lo : (c == 1) ? hi : roots[c - 2];
      if (t < lo || t > hi) continue;  // also drops NaN times
      const double m = vals[c];
      if (mx_v < m) {  // std::max keeps the first on ties
        mx_v = m;
        mx_t = t;
      }
      if (m < mn_v) {
        mn_v = m;
        mn_t = t;
      }
      if (w < p.max_cand) {
        if (p.cand_time) p.cand_time[at<AOS>((size_t)seg * p.max_cand + w, rec_k, Bsz, (size_t)b)] = t;
        if (p.cand_value) p.cand_value[at<AOS>((size_t)seg * p.max_cand + w, rec_k, Bsz, (size_t)b)] = m;
      }
      ++w;
    }
    uint32_t st = (uint32_t)s_st[q];
    if ((p.cand_time || p.cand_value) && w > p.max_cand) st |= 8u;  // MTG_ST_TRUNCATED
    if (p.n_cand) p.n_cand[at<AOS>((size_t)seg, (size_t)K, Bsz, (size_t)b)] = w;
    if (p.seg_out) {
      double* o = p.seg_out + ((size_t)local * K + seg) * 4;
      o[0] = mn_t;
      o[1] = mn_v;
      o[2] = mx_t;
      o[3] = mx_v;
      p.seg_status[(size_t)local * K + seg] = st;
    }
  }
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
  if (p.status) p.status[b] = p.soft_accumulate && p.soft_cost ? (p.status[b] | st) : st;
  if (p.soft_violation) p.soft_violation[b] = mx_v - p.soft_limit;
  if (p.soft_cost) {
    const double c = fmin(p.soft_max, exp((mx_v - p.soft_limit) / p.soft_limit * p.soft_weight));
    p.soft_cost[b] = p.soft_accumulate ? p.soft_cost[b] + c : c;
  }
}

}  // namespace mtg
#endif

// E5 (analytic) / E6 / R1 — extrema of the derivative magnitude |p^(d)(t)| per segment and per
// trajectory, the candidate lists behind them, and the real roots of a polynomial in an interval.
//
// Replaces (reference): Segment::computeMinMaxMagnitudeCandidateTimes segment.cpp:82-133
// (g = sum_dim conv(delta, delta'), polynomial.cpp:163-181; one dimension: roots of p^(d+1)),
// Polynomial::computeMinMaxCandidates / selectMinMaxCandidatesFromRoots polynomial.cpp:32-83,
// findRootsJenkinsTraub + rpoly_ak1 rpoly_ak1.cpp:57-937, the candidate evaluation
// segment.cpp:135-184 and Trajectory::computeMinMaxMagnitude trajectory.cpp:184-220.
//
// Root finding. The reference runs Jenkins-Traub for ALL complex roots and keeps the real ones
// inside [t_start, t_end] (polynomial.cpp:46-60). Only those are ever consumed, so this kernel
// isolates exactly them with the derivative chain: the roots of P^(k+1) split the interval into
// pieces on which P^(k) is monotone; every piece whose end values differ in sign holds exactly one
// root of P^(k), refined by a bracketed Newton iteration (bisection when a Newton step leaves the
// bracket). Going from the linear P^(n-1) down to P^(0) = g yields every real root of g in the
// interval that is a sign change — i.e. every extremum of the magnitude; roots of even
// multiplicity are inflections of the magnitude and cannot be its minimum or maximum. All loops
// are bounded and there is no data-dependent recursion.
//
// Mapping (warp-cooperative). One warp owns kExG = 16 root problems ((trajectory, segment) pairs,
// contiguous in memory in both layouts). Everything a problem needs between levels — g, the
// current level polynomial, the roots of the previous and of the current level — lives in
// shared memory (odd strides: the 16 problems of a warp sit in 16 different bank pairs), nothing
// in local memory. Per level the warp runs three flat, lane-parallel passes over ALL of its
// problems at once:
//   build   lane = (problem, coefficient): the level polynomial from g and the base table;
//   scan    lane = (problem, piece of the partition): both end values of the piece in one Horner
//           sweep (one shared-memory load feeds two FMA chains), sign test, ordered compaction of
//           the brackets with ballots (roots stay sorted per problem);
//   refine  lane = bracket: bracketed Newton with the value and the slope from ONE coefficient
//           stream (p and p' by the coupled Horner recurrence).
// So the lanes of a warp share the work of 16 problems instead of each waiting for the slowest
// of 32 (root counts and iteration counts differ from problem to problem, their sums over 16
// problems hardly do). The candidates [t_start, t_end, roots...] are then evaluated one per lane.
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
constexpr int kExG = 16;     // root problems per warp (the item -> problem search below is written for 16)
static_assert(kExG == 16, "the prefix search of extrema_warp_kernel assumes 16 problems per warp");
constexpr int kExWarps = 8;  // warps per CTA

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
  int S;       // stride (doubles) of the per-problem arrays g, pk, rA, rB
  int nd;      // coefficients of p^(d)
  int ndim;    // dimensions taking part
  size_t warp_bytes, cta_bytes;
};

__host__ __device__ inline ExtremaPlan extrema_plan(int N, int D, int derivative, int dim_mask, int raw) {
  ExtremaPlan pl;
  int ndim = 0;
  for (int q = 0; q < D; ++q) ndim += (dim_mask >> q) & 1;
  pl.ndim = ndim;
  pl.nd = raw ? N : N - derivative;
  pl.len = raw ? N : (ndim > 1 ? 2 * pl.nd - 2 : pl.nd - 1);
  if (pl.len < 1) pl.len = 1;
  // every per-problem array holds <= len + 1 doubles (len coefficients; len - 1 roots + 2 end points);
  // [g | pk] together must also hold the staged derivative coefficients (D * nd) for the candidates
  // (and [pk | rA | rB] the same for building g)
  int S = (pl.len + 1) | 1;
  const int need = raw ? 0 : ((D * pl.nd + 1) / 2) | 1;
  if (need > S) S = need;
  pl.S = S;
  // doubles: 4 arrays x G x S + lo/hi[G];  ints: n, na, cnt, par, st [G], off[G + 1], next;  uint16 ent[G * len]
  size_t bytes = (size_t)(4 * kExG * S + 2 * kExG) * sizeof(double) + (size_t)(5 * kExG + kExG + 2) * sizeof(int) +
                 (size_t)kExG * pl.len * sizeof(uint16_t);
  pl.warp_bytes = (bytes + 15) & ~(size_t)15;
  pl.cta_bytes = pl.warp_bytes * kExWarps + (size_t)MTG_BASE_LD * MTG_BASE_LD * sizeof(double);
  return pl;
}

}  // namespace mtg
struct mtg_ctx;
namespace mtg {
// extrema.cu: chunked launch of extrema_warp_kernel (+ extrema_reduce_kernel when per-trajectory outputs are wanted)
int launch_extrema(mtg_ctx* ctx, bool aos, const ExtremaParams& p, cudaStream_t s);

__device__ __forceinline__ unsigned lanes_lt(int lane) { return (1u << lane) - 1u; }

template <bool AOS>
__global__ void __launch_bounds__(kExWarps * 32) extrema_warp_kernel(const ExtremaParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ExtremaPlan pl = extrema_plan(p.N, p.D, p.derivative, p.dim_mask, p.raw);
  const int len = pl.len, S = pl.S, nd = pl.nd, G = kExG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr unsigned FULL = 0xffffffffu;

  // ---- CTA-wide: the base table B(k, j) = j!/(j-k)! in shared memory (lanes read different rows)
  double* s_base = reinterpret_cast<double*>(smem_raw + pl.warp_bytes * kExWarps);
  for (int i = threadIdx.x; i < MTG_BASE_LD * MTG_BASE_LD; i += blockDim.x) s_base[i] = c_base.base[i];
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
  double* s_g = reinterpret_cast<double*>(base_ptr);
  double* s_pk = s_g + G * S;
  double* s_ra = s_pk + G * S;
  double* s_rb = s_ra + G * S;
  double* s_lo = s_rb + G * S;
  double* s_hi = s_lo + G;
  int* s_n = reinterpret_cast<int*>(s_hi + G);
  int* s_na = s_n + G;
  int* s_cnt = s_na + G;
  int* s_par = s_cnt + G;
  int* s_st = s_par + G;
  int* s_off = s_st + G;  // G + 1
  int* s_next = s_off + G + 1;
  uint16_t* s_ent = reinterpret_cast<uint16_t*>(s_next + 1);

  const size_t Bsz = (size_t)p.B;
  const size_t rec_c = p.raw ? (size_t)N : (size_t)K * D * N;
  const int rec_one = p.raw ? N : D * N;  // coefficients of one problem

  // ---- interval of every problem
  if (lane < G) {
    double lo = 0.0, hi = 1.0;
    if (lane < np) {
      const int b = p.b0 + prob_local(lane), seg = prob_seg(lane);
      const size_t o = at<AOS>((size_t)seg, (size_t)K, Bsz, (size_t)b);
      lo = p.t_lo ? p.t_lo[o] : 0.0;
      hi = p.t_hi ? p.t_hi[o] : p.seg_times[o];
    }
    s_lo[lane] = lo;
    s_hi[lane] = hi;
    s_st[lane] = 0;
    s_par[lane] = 0;
    s_na[lane] = 0;
  }

  // Stages the derivative coefficients delta[dim][j] = B(d, j+d) c[j+d] (polynomial.h:99-113) of all
  // problems at `dst` (stride sd per problem, dims not in dim_mask zeroed), coalesced in both layouts.
  // In raw mode the record is the polynomial itself.
  auto stage_delta = [&](double* dst, int sd) {
    const int n_el = np * rec_one;
    // element e = lane + 32 k walks memory in order; (q, r) follow it with counters (no division per element)
    int q = AOS ? lane / rec_one : 0, r = AOS ? lane - q * rec_one : lane / np;
    int qs = AOS ? 0 : lane - r * np;  // SoA: problem within the row of 16
    for (int e = lane; e < n_el; e += 32) {
      const int qq = AOS ? q : qs;
      const int b = p.b0 + prob_local(qq), seg = prob_seg(qq);
      if (p.raw) {
        dst[qq * sd + r] = p.coeffs[at<AOS>((size_t)r, rec_c, Bsz, (size_t)b)];
      } else {
        int dim = 0, jj = r;
        while (jj >= N) {
          jj -= N;
          ++dim;
        }
        if (jj >= d) {
          const double c = p.coeffs[at<AOS>((size_t)(seg * D + dim) * N + jj, rec_c, Bsz, (size_t)b)];
          dst[qq * sd + dim * nd + (jj - d)] = ((p.dim_mask >> dim) & 1) ? s_base[d * MTG_BASE_LD + jj] * c : 0.0;
        }
      }
      if (AOS) {
        r += 32;
        while (r >= rec_one) {
          r -= rec_one;
          ++q;
        }
      } else {
        qs += 32;
        while (qs >= np) {
          qs -= np;
          ++r;
        }
      }
    }
  };

  // ---- g: the polynomial whose real roots in [lo, hi] are the candidate times
  {
    double* s_delta = s_pk;  // [pk | rA | rB] holds D * nd doubles per problem at this point
    const int sd = p.raw ? N : D * nd;
    stage_delta(s_delta, sd);
    __syncwarp();
    // lane = (problem q, coefficient m = 2 k + half): 16 problems x 2 lanes, conflict-free odd strides
    const int q = lane & 15;
    if (q < np) {
      const double* dl = s_delta + q * sd;
      for (int m = lane >> 4; m < len; m += 2) {
        double acc = 0.0;
        if (p.raw) {
          acc = dl[m];
        } else if (pl.ndim > 1) {
          // sum_dim conv(delta, delta'), delta'[j] = (j+1) delta[j+1]   (segment.cpp:93-115)
          const int i0 = max(0, m - (nd - 2)), i1 = min(m, nd - 1);
          for (int dim = 0; dim < D; ++dim) {
            double f = (double)(m - i0 + 1);
            for (int i = i0; i <= i1; ++i) {
              acc = fma(dl[dim * nd + i], f * dl[dim * nd + m - i + 1], acc);
              f -= 1.0;
            }
          }
        } else {
          // one dimension: roots of p^(d+1)   (segment.cpp:124-131); the other dimensions are staged as zeros
          for (int dim = 0; dim < D; ++dim) acc += (double)(m + 1) * dl[dim * nd + m + 1];
        }
        s_g[q * S + m] = acc;
      }
    }
    __syncwarp();
  }
  // strip zero leading coefficients (findLastNonZeroCoeff, rpoly_ak1.cpp:57-68)
  int nmax = 0;
  {
    int n = -1;
    if (lane < np) {
      n = len - 1;
      while (n >= 0 && !(fabs(s_g[lane * S + n]) >= 2.2250738585072014e-308)) --n;
    }
    if (lane < G) s_n[lane] = n;
    nmax = n;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) nmax = max(nmax, __shfl_xor_sync(FULL, nmax, m));
    __syncwarp();
  }

  // ---- the derivative chain, level by level: deg = degree of the level polynomial g^(n - deg)
  for (int deg = 1; deg <= nmax; ++deg) {
    // build: pk[q][j] = B(k, j + k) g[q][j + k], k = n_q - deg, for the problems still in the chain
    {
      const int q = lane & 15;  // lane = (problem, coefficient parity)
      const int k = q < np ? s_n[q] - deg : -1;
      if (k >= 0) {
        const double* bk = s_base + k * MTG_BASE_LD + k;
        const double* gk = s_g + q * S + k;
        double* pq = s_pk + q * S;
        for (int j = lane >> 4; j <= deg; j += 2) pq[j] = bk[j] * gk[j];
      }
    }
    // pieces of the partition per problem (exclusive prefix over the problems)
    {
      int cnt = 0;
      if (lane < np && s_n[lane] >= deg) cnt = s_na[lane] + 1;
      int incl = cnt;
#pragma unroll
      for (int m = 1; m < 32; m <<= 1) {
        const int o = __shfl_up_sync(FULL, incl, m);
        if (lane >= m) incl += o;
      }
      if (lane < G) {
        s_off[lane] = incl - cnt;
        s_cnt[lane] = 0;
      }
      if (lane == G - 1) s_off[G] = incl;
    }
    __syncwarp();
    const int total = s_off[G];
    int nent = 0;
    // scan: lane = piece [u, v] of one problem's partition
    for (int it0 = 0; it0 < total; it0 += 32) {
      const int item = it0 + lane;
      const bool live = item < total;
      int q = 0;
      if (live) {
        if (item >= s_off[8]) q = 8;
        if (item >= s_off[q + 4]) q += 4;
        if (item >= s_off[q + 2]) q += 2;
        if (item >= s_off[q + 1]) q += 1;
      }
      const int qi = item - s_off[q];
      const int na = s_na[q];
      const double* rprev = (s_par[q] ? s_rb : s_ra) + q * S;
      double* rcur = (s_par[q] ? s_ra : s_rb) + q * S;
      const double lo = s_lo[q], hi = s_hi[q];
      double u = lo, v = hi;
      if (live) {
        if (qi > 0) u = rprev[qi - 1];
        if (qi < na) v = rprev[qi];
      }
      const bool valid = live && (v > u);
      double fu = 0.0, fv = 0.0;
      if (valid) {
        const double* pk = s_pk + q * S;
        fu = fv = pk[deg];
#pragma unroll 4
        for (int j = deg - 1; j >= 0; --j) {
          const double c = pk[j];
          fu = fma(fu, u, c);
          fv = fma(fv, v, c);
        }
      }
      // what this piece contributes (in root order): a root exactly on its left end, or a bracket;
      // and, for the piece that ends at hi, a root exactly on hi
      const bool exact_u = valid && fu == 0.0;
      const bool bracket = valid && fu != 0.0 && fv != 0.0 && ((fu < 0.0) != (fv < 0.0));
      const bool exact_v = valid && fv == 0.0 && v == hi;
      const bool e1 = exact_u || bracket;
      // lanes of the same problem are contiguous: [first, last] within this round
      const int first = max(0, s_off[q] - it0), last = min(31, s_off[q + 1] - 1 - it0);
      const unsigned same = live ? ((last >= 31 ? FULL : ((1u << (last + 1)) - 1u)) & ~lanes_lt(first)) : 0u;
      const unsigned m1 = __ballot_sync(FULL, e1);
      const unsigned mb = __ballot_sync(FULL, bracket);
      const int before = live ? s_cnt[q] : 0;
      __syncwarp();
      const int slot = before + __popc(m1 & same & lanes_lt(lane));
      if (e1 && (m1 & same & lanes_lt(lane)) == 0u) s_cnt[q] = before + __popc(m1 & same);  // first emitter of q
      if (exact_u) rcur[slot] = u;
      if (bracket) {
        // first iterate: the secant point (the midpoint if it degenerates), parked where the root will go
        double t = u - fu * ((v - u) / (fv - fu));
        if (!(t > u && t < v)) t = 0.5 * (u + v);
        rcur[slot] = t;
        s_ent[nent + __popc(mb & lanes_lt(lane))] =
            (uint16_t)(q | (qi << 4) | (slot << 9) | ((fu < 0.0) ? (1 << 14) : 0));
      }
      nent += __popc(mb);
      __syncwarp();
      if (exact_v) {  // at most one piece per problem ends at hi
        const int s2 = s_cnt[q];
        rcur[s2] = v;
        s_cnt[q] = s2 + 1;
      }
      __syncwarp();
    }
    // refine: lane = bracket. ONE flat loop whose trip is one bracketed-Newton iteration of whichever bracket
    // the lane is working on; a lane that finishes a bracket draws the next one from a shared counter, so the
    // warp waits for the largest SUM of iterations per lane, not for the slowest bracket of every round.
    if (lane == 0) *s_next = 32;
    __syncwarp();
    {
      int e = lane;
      bool have = false;
      int q = 0, slot = 0, it = 0;
      bool fa_neg = false;
      double a = 0.0, bb = 0.0, t = 0.0, tol = 0.0, wmin = 0.0;
      const double* pk = s_pk;
      double* rcur = s_rb;
      for (;;) {
        if (!have) {
          if (e >= nent) break;
          const unsigned en = s_ent[e];
          q = en & 15;
          const int qi = (en >> 4) & 31;
          slot = (en >> 9) & 31;
          fa_neg = (en >> 14) & 1;
          const int na = s_na[q];
          const double* rprev = (s_par[q] ? s_rb : s_ra) + q * S;
          rcur = (s_par[q] ? s_ra : s_rb) + q * S;
          a = qi > 0 ? rprev[qi - 1] : s_lo[q];
          bb = qi < na ? rprev[qi] : s_hi[q];
          // roots of the upper levels only PARTITION the interval for the level below: 1e-9 of its length is
          // plenty (a partition point off by delta can only hide a root pair closer than delta, i.e. a bump of
          // the magnitude of relative height ~delta^2); the roots of g itself (deg = n) go to 1e-15 of the
          // length (a few ulp of a mid-interval time)
          tol = ((s_n[q] > deg) ? 1e-9 : 1e-15) * (s_hi[q] - s_lo[q]);
          // bracket narrower than tol or than fp64 can resolve anywhere in the interval
          wmin = fmax(tol, 4.5e-16 * fmax(fabs(s_lo[q]), fabs(s_hi[q])));
          pk = s_pk + q * S;
          t = rcur[slot];
          it = 0;
          have = true;
        }
        // value, slope and the running rounding-error bound of the value from one coefficient stream:
        // p and p' by the coupled Horner recurrence, err = sum |c_j| |t|^j
        double ft = pk[deg], dft = 0.0, err = fabs(ft);
        const double at = fabs(t);
#pragma unroll 4
        for (int j = deg - 1; j >= 0; --j) {
          const double c = pk[j];
          dft = fma(dft, t, ft);
          ft = fma(ft, t, c);
          err = fma(err, at, fabs(c));
        }
        double res = t;
        // |value| inside its own rounding noise: the root is located as well as fp64 can tell. (This is what
        // ends the iteration on the numerically multiple roots of rest-to-rest ends, where Newton converges
        // only linearly and every further digit is noise anyway.)
        bool fin = fabs(ft) <= (double)(2 * deg + 2) * 1.1102230246251565e-16 * err;
        if (!fin) {
          if ((ft < 0.0) == fa_neg)
            a = t;
          else
            bb = t;
          const double width = bb - a;
          if (!(width > wmin)) {
            fin = true;
          } else {
            // Newton step with a cheap reciprocal (rcp.approx.f64: ~20 bits over the whole fp64 exponent range,
            // + one Newton step: ~1e-12 relative — the step only has to land inside the bracket; anything
            // else, incl. NaN / infinity from a vanishing slope, bisects)
            double r;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(dft));
            r = r * fma(-dft, r, 2.0);
            double tn = fma(-ft, r, t);
            if (!(tn > a && tn < bb)) tn = 0.5 * (a + bb);
            if (!(fabs(tn - t) > tol)) {
              fin = true;
              res = tn;
            } else {
              t = tn;
              if (++it >= kRootIters) {
                atomicOr(&s_st[q], 16);  // MTG_ST_NO_CONVERGENCE (reference: rpoly returns partial roots, RPOLY_C:372-377)
                fin = true;
                res = t;
              }
            }
          }
        }
        if (fin) {
          rcur[slot] = res;
          have = false;
          e = atomicAdd(s_next, 1);
        }
      }
    }
    __syncwarp();
    // the problems of this level: current roots become the partition of the next level
    if (lane < np && s_n[lane] >= deg) {
      s_na[lane] = s_cnt[lane];
      s_par[lane] ^= 1;
    }
    __syncwarp();
  }
  // roots of g of problem q: (par ? rB : rA)[0 .. na) ascending; constant polynomial: none (rpoly_ak1.cpp:76-80)

  // ---- raw mode: the roots are the result
  if (p.raw) {
    if (lane < np) {
      const int q = lane;
      const int b = p.b0 + prob_local(q), seg = prob_seg(q);
      const double* roots = (s_par[q] ? s_rb : s_ra) + q * S;
      const int na = s_n[q] >= 1 ? s_na[q] : 0;
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
  double* s_delta = s_g;  // [g | pk] are free now
  const int sd = D * nd;
  stage_delta(s_delta, sd);
  {
    int cnt = 0;
    if (lane < np) {
      if (s_n[lane] < 1) s_na[lane] = 0;
      cnt = s_na[lane] + 2;
    }
    int incl = cnt;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
      const int o = __shfl_up_sync(FULL, incl, m);
      if (lane >= m) incl += o;
    }
    if (lane < G) s_off[lane] = incl - cnt;
    if (lane == G - 1) s_off[G] = incl;
  }
  __syncwarp();
  {
    const int total = s_off[G];
    for (int item = lane; item < total; item += 32) {
      int q = 0;
      if (item >= s_off[8]) q = 8;
      if (item >= s_off[q + 4]) q += 4;
      if (item >= s_off[q + 2]) q += 2;
      if (item >= s_off[q + 1]) q += 1;
      const int c = item - s_off[q];
      const double* roots = (s_par[q] ? s_rb : s_ra) + q * S;
      double* vals = (s_par[q] ? s_ra : s_rb) + q * S;  // the other root buffer is free
      const double t = (c == 0) ? s_lo[q] : (c == 1) ? s_hi[q] : roots[c - 2];
      const double* dl = s_delta + q * sd;
      double m2 = 0.0;
      for (int dim = 0; dim < D; ++dim) {
        double r = 0.0;
        for (int j = nd - 1; j >= 0; --j) r = fma(r, t, dl[dim * nd + j]);
        m2 = fma(r, r, m2);
      }
      vals[c] = sqrt(m2);
    }
  }
  __syncwarp();
  if (lane < np) {
    const int q = lane;
    const int local = prob_local(q), seg = prob_seg(q), b = p.b0 + local;
    const double* roots = (s_par[q] ? s_rb : s_ra) + q * S;
    const double* vals = (s_par[q] ? s_ra : s_rb) + q * S;
    const double lo = s_lo[q], hi = s_hi[q];
    const int na = s_na[q];
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

// libmtg_cuda.so — N1: the batched non-linear objective over the free endpoint derivatives and a
// projected-gradient driver on top of it (the part of PolynomialOptimizationNonLinear that is plain
// data-parallel arithmetic; NLOPT itself, the octree and the collision term are not reproduced).
//
// Replaces (reference, impl/polynomial_optimization_nonlinear_impl.h = NL_I):
//   getCostAndGradientDerivative        NL_I:1537-1606   J_d = sum_dim [d_f; d_p]^T R [d_f; d_p] (no 1/2) and
//                                                        grad_{d_p} = 2 R_pf d_f + 2 R_pp d_p
//   getCostAndGradientSoftConstraints   NL_I:2365-2423   J_sc and its CENTRAL finite-difference gradient: for every
//     (+ ...Simple, forward, :2425-2490)                  free derivative, setFreeConstraints(d_p -+ increment) and
//                                                        evaluateMaximumMagnitudeAsSoftConstraint (:2735-2766)
//   objectiveFunctionFreeConstraints[AndCollision]  NL_I:1024-1113, 1115-1284 (w_d J_d + w_sc J_sc, no collision term)
//   setFreeEndpointDerivativeHardConstraints        NL_I:2858-2905 (|d_p| <= the v / a limits)
//
// How the reference's work collapses here:
//  * R = C^T blockdiag(H_i) C, so J_d = sum_i q_i and dJ_d/d(v, k) = [dq_{v-1}/d(end k)] + [dq_v/d(start k)]:
//    two rows of H_i applied to the segment's own 2h endpoint derivatives — no 55 x 55 R is ever formed
//    (the reference builds the dense R with getR on every evaluation, NL_I:1552-1554). One thread per trajectory.
//  * a change of ONE free derivative (v, k) of ONE dimension only changes the polynomials of segments v-1 and v,
//    linearly: c_new = c -+ increment T^(k-j) A(1)^-1[j][col]. The 2 D (K-1)(h-1) perturbed trajectories of the
//    reference's finite-difference loop (each a full setFreeConstraints + rpoly over all K segments) become a
//    batch of 2-SEGMENT root problems for the warp-cooperative extrema kernel; the maxima of the K-2 untouched
//    segments are taken from the nominal evaluation.
#include <cmath>

#include "host_common.h"
#include "cost_fd.cuh"
#include "extrema.cuh"

MTG_REGISTER_TABLES()
MTG_REGISTER_BASE()

using namespace mtg;

namespace {

// ------------------------------------------------------------------ J_d and its analytic gradient
struct CostDerivParams {
  CostFdParams in;                 // positions, end_derivatives, seg_times, free_constraints, B, b0, nb, K, derivative
  double* __restrict__ J_d;        // [B] or nullptr
  double* __restrict__ grad;       // elem ((dim*(K-1) + v-1)*(h-1) + k-1), rec D*(K-1)*(h-1); or nullptr
  double* __restrict__ diag;       // elem ((v-1)*(h-1) + k-1), rec (K-1)*(h-1): diagonal of 2 R_pp; or nullptr
};

template <int HN, int D, bool AOS>
__global__ void __launch_bounds__(128) cost_derivative_kernel(const CostDerivParams q) {
  constexpr int NF = HN - 1, N = 2 * HN;
  const CostFdParams& p = q.in;
  const int local = blockIdx.x * blockDim.x + threadIdx.x;
  if (local >= p.nb) return;
  const int b = p.b0 + local;
  const size_t B = (size_t)p.B;
  const int K = p.K, der = p.derivative;
  const size_t rec_pos = (size_t)(K + 1) * D, rec_end = (size_t)2 * NF * D, rec_free = (size_t)D * (K - 1) * NF;
  const size_t rec_diag = (size_t)(K - 1) * NF;
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
  const int nq = N - der;
  double ds[D][HN], de[D][HN];
  double ge_prev[D][HN];   // d q_{i-1} / d(end derivative m) of the previous segment
  double dg_prev[HN];      // its diagonal terms
  vertex(0, ds);
  double J = 0.0;
  for (int i = 0; i < K; ++i) {
    double T = p.seg_times[at<AOS>((size_t)i, (size_t)K, B, b)];
    if (!(T > 0.0) || !(T < 1.7e308)) {  // LIN_I:296 CHECK_GT(segment_time, 0)
      st |= 1u;
      T = 1.0;
    }
    vertex(i + 1, de);
    double tp[HN];
    tp[0] = 1.0;
#pragma unroll
    for (int m = 1; m < HN; ++m) tp[m] = tp[m - 1] * T;
    double s = 1.0;  // T^(1-2d)
    {
      const int e0 = 1 - 2 * der;
      const double u = 1.0 / T;
      for (int r = 0; r < (e0 >= 0 ? e0 : -e0); ++r) s *= (e0 >= 0 ? T : u);
    }
    double quad = 0.0;
    double gs[D][HN], ge[D][HN];
#pragma unroll
    for (int dim = 0; dim < D; ++dim) {
      double dhs[HN], dhe[HN];
      const double dlt = de[dim][0] - ds[dim][0];
#pragma unroll
      for (int m = 1; m < HN; ++m) {
        dhs[m] = tp[m] * ds[dim][m];
        dhe[m] = tp[m] * de[dim][m];
        gs[dim][m] = 0.0;
        ge[dim][m] = 0.0;
      }
#pragma unroll
      for (int r = 0; r < N; ++r) {
        if (r < nq) {
          double w = MTG_W(r, HN) * dlt;
          if (der == 0) w = fma(MTG_W(r, 0) + MTG_W(r, HN), ds[dim][0], w);
#pragma unroll
          for (int m = 1; m < HN; ++m) {
            w = fma(MTG_W(r, m), dhs[m], w);
            w = fma(MTG_W(r, HN + m), dhe[m], w);
          }
          quad = fma(w, w, quad);
#pragma unroll
          for (int m = 1; m < HN; ++m) {
            gs[dim][m] = fma(w, MTG_W(r, m), gs[dim][m]);
            ge[dim][m] = fma(w, MTG_W(r, HN + m), ge[dim][m]);
          }
        }
      }
#pragma unroll
      for (int m = 1; m < HN; ++m) {  // d q / d d_m = 2 s T^m (W^T w)_m
        gs[dim][m] *= 2.0 * s * tp[m];
        ge[dim][m] *= 2.0 * s * tp[m];
      }
    }
    J = fma(quad, s, J);
    double dgs[HN], dge[HN];  // diagonal of 2 H_i at the start / end derivative m
#pragma unroll
    for (int m = 1; m < HN; ++m) {
      dgs[m] = 2.0 * s * tp[m] * tp[m] * MTG_H1(m, m);
      dge[m] = 2.0 * s * tp[m] * tp[m] * MTG_H1(HN + m, HN + m);
    }
    if (i >= 1) {  // vertex i is free: both of its segments are known now
#pragma unroll
      for (int m = 1; m < HN; ++m) {
        if (q.grad) {
#pragma unroll
          for (int dim = 0; dim < D; ++dim)
            q.grad[at<AOS>((size_t)(dim * (K - 1) + (i - 1)) * NF + (m - 1), rec_free, B, b)] = ge_prev[dim][m] + gs[dim][m];
        }
        if (q.diag) q.diag[at<AOS>((size_t)(i - 1) * NF + (m - 1), rec_diag, B, b)] = dg_prev[m] + dgs[m];
      }
    }
#pragma unroll
    for (int m = 1; m < HN; ++m) {
      dg_prev[m] = dge[m];
#pragma unroll
      for (int dim = 0; dim < D; ++dim) ge_prev[dim][m] = ge[dim][m];
    }
#pragma unroll
    for (int dim = 0; dim < D; ++dim)
#pragma unroll
      for (int m = 0; m < HN; ++m) ds[dim][m] = de[dim][m];
  }
  if (q.J_d) q.J_d[b] = J;
  if (p.status) p.status[b] = st;
}

template <int HN, int D, bool AOS>
int launch_cdv_t(mtg_ctx* ctx, const CostDerivParams& p, cudaStream_t s) {
  const int block = 128, grid = (p.in.nb + block - 1) / block;
  if (grid == 0) return MTG_OK;
  cost_derivative_kernel<HN, D, AOS><<<grid, block, 0, s>>>(p);
  ++ctx->launches;
  MTG_CUDA_TRY(cudaGetLastError());
  return MTG_OK;
}
template <int HN, bool AOS>
int launch_cdv_d(mtg_ctx* ctx, int D, const CostDerivParams& p, cudaStream_t s) {
  switch (D) {
    case 1: return launch_cdv_t<HN, 1, AOS>(ctx, p, s);
    case 2: return launch_cdv_t<HN, 2, AOS>(ctx, p, s);
    case 3: return launch_cdv_t<HN, 3, AOS>(ctx, p, s);
    case 4: return launch_cdv_t<HN, 4, AOS>(ctx, p, s);
  }
  return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "D must be 1..4");
}
template <bool AOS>
int launch_cdv_n(mtg_ctx* ctx, int N, int D, const CostDerivParams& p, cudaStream_t s) {
  switch (N) {
    case 4: return launch_cdv_d<2, AOS>(ctx, D, p, s);
    case 6: return launch_cdv_d<3, AOS>(ctx, D, p, s);
    case 8: return launch_cdv_d<4, AOS>(ctx, D, p, s);
    case 10: return launch_cdv_d<5, AOS>(ctx, D, p, s);
    case 12: return launch_cdv_d<6, AOS>(ctx, D, p, s);
  }
  return fail(ctx, MTG_ERR_UNSUPPORTED, "supported N: {4,6,8,10,12}");
}

// ------------------------------------------------------------------ soft-constraint gradient
// Perturbed 2-segment trajectories. Variable q = (dim, v, k) (the order of free_constraints), sign s in {-, +}:
// item = ((local * Q + q) * 2 + s) holds segments v-1 and v with d_p[dim][v][k] -+ increment, as an AoS batch
// of K' = 2 segment trajectories: coeffs [item][2][D][N], times [item][2].
struct PerturbParams {
  const double* __restrict__ coeffs;     // nominal, elem ((i*D + dim)*N + j), rec K*D*N
  const double* __restrict__ seg_times;  // elem i, rec K
  double* __restrict__ pc;               // [nb*Q*2][2][D][N]
  double* __restrict__ pt;               // [nb*Q*2][2]
  double increment;
  int B, b0, nb, K, D, N;
};

template <bool AOS>
__global__ void __launch_bounds__(256) perturb_segments_kernel(const PerturbParams p) {
  const int K = p.K, D = p.D, N = p.N, h = N / 2, NF = h - 1;
  const int Q = D * (K - 1) * NF;
  const size_t per_item = (size_t)2 * D * N;
  const size_t total = (size_t)p.nb * Q * 2 * per_item;
  const size_t B = (size_t)p.B, rec_c = (size_t)K * D * N;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
    const size_t item = g / per_item;
    const int e = (int)(g - item * per_item);
    const int j = e % N, dim = (e / N) % D, half = e / (N * D);  // half 0: segment v-1, 1: segment v
    const int sgn = (int)(item & 1);
    const size_t lq = item >> 1;
    const int qv = (int)(lq % Q);
    const int local = (int)(lq / Q);
    const int k = qv % NF + 1, v = (qv / NF) % (K - 1) + 1, pdim = qv / (NF * (K - 1));
    const int b = p.b0 + local;
    const int seg = v - 1 + half;
    double c = p.coeffs[at<AOS>((size_t)(seg * D + dim) * N + j, rec_c, B, (size_t)b)];
    const double T = p.seg_times[at<AOS>((size_t)seg, (size_t)K, B, (size_t)b)];
    if (dim == pdim) {
      // c = A(T)^-1 d,  A(T)^-1[j][col] = T^-j A(1)^-1[j][col] T^alpha(col):  column h + k of segment v-1 (its
      // end), column k of segment v (its start); alpha = k for both
      const int col = half == 0 ? h + k : k;
      double f = 1.0;
      const int e0 = k - j;
      const double u = 1.0 / T;
      for (int r = 0; r < (e0 >= 0 ? e0 : -e0); ++r) f *= (e0 >= 0 ? T : u);
      const double dc = p.increment * f * c_tab.Ainv1[j * MTG_TAB_LD + col];
      c = sgn ? c + dc : c - dc;
    }
    p.pc[g] = c;
    if (e == 0) {
      p.pt[item * 2] = p.seg_times[at<AOS>((size_t)(v - 1), (size_t)K, B, (size_t)b)];
      p.pt[item * 2 + 1] = p.seg_times[at<AOS>((size_t)v, (size_t)K, B, (size_t)b)];
    }
  }
}

constexpr int kMaxSoft = 4;  // inequality constraints (the reference uses two: v_max, a_max)
struct SoftCombineParams {
  const double* nominal[kMaxSoft];    // per constraint: nominal per-segment maxima, AoS [nb][K]
  const double* perturbed[kMaxSoft];  // per constraint: maxima of the perturbed 2-segment items, AoS [nb*Q*2][2]
  double limit[kMaxSoft];
  double* __restrict__ J_sc;          // [B] or nullptr
  double* __restrict__ grad;          // elem q, rec Q = D (K-1)(h-1); or nullptr
  double weight, max_cost, increment;
  int central;
  int n_con, B, b0, nb, K, Q, NF;
  int cost_only;  // 1: J_sc only (Q = 1, no perturbed maxima)
};

__device__ __forceinline__ double soft_cost(double mx, double limit, double w, double cap) {
  return fmin(cap, exp((mx - limit) / limit * w));  // NL_I:2753-2756
}

template <bool AOS>
__global__ void __launch_bounds__(128) soft_combine_kernel(const SoftCombineParams p) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (size_t)p.nb * p.Q) return;
  const int local = (int)(g / p.Q), q = (int)(g % p.Q);
  const int K = p.K;
  const int v = p.cost_only ? 1 : (q / p.NF) % (K - 1) + 1;
  double c_nom = 0.0, c_lo = 0.0, c_hi = 0.0;
  for (int c = 0; c < p.n_con; ++c) {
    const double* nm = p.nominal[c] + (size_t)local * K;
    double rest = -1.7976931348623157e308, all = -1.7976931348623157e308;
    for (int s = 0; s < K; ++s) {
      const double m = nm[s];
      all = fmax(all, m);
      if (s != v - 1 && s != v) rest = fmax(rest, m);
    }
    c_nom += soft_cost(all, p.limit[c], p.weight, p.max_cost);
    if (p.cost_only) continue;
    const double* pm = p.perturbed[c] + ((size_t)local * p.Q + q) * 4;
    c_lo += soft_cost(fmax(rest, fmax(pm[0], pm[1])), p.limit[c], p.weight, p.max_cost);
    c_hi += soft_cost(fmax(rest, fmax(pm[2], pm[3])), p.limit[c], p.weight, p.max_cost);
  }
  const int b = p.b0 + local;
  if (p.grad)
    p.grad[at<AOS>((size_t)q, (size_t)p.Q, (size_t)p.B, (size_t)b)] =
        p.central ? (c_hi - c_lo) / (2.0 * p.increment) : (c_hi - c_nom) / p.increment;
  if (p.J_sc && q == 0) p.J_sc[b] = c_nom;
}

// ------------------------------------------------------------------ projected gradient with step rejection
// State per trajectory: the last ACCEPTED point x_prev with its search direction g_prev and objective f_prev,
// and a step length. Every pass evaluates the objective and gradients at the current trial point; the trial is
// accepted iff f did not increase (a NaN never is), otherwise the step is halved and the next trial starts again
// from x_prev along g_prev. No extra objective evaluations are needed.
struct DescentParams {
  double* __restrict__ x;              // trial point: free constraints, elem ((dim*(K-1) + v-1)*NF + k-1), rec Q
  double* __restrict__ x_prev;         // last accepted point, same layout
  double* __restrict__ g_prev;         // its (preconditioned, weighted) direction, same layout
  const double* __restrict__ grad_d;   // same layout
  const double* __restrict__ grad_sc;  // same layout or nullptr
  const double* __restrict__ diag;     // elem ((v-1)*NF + k-1), rec (K-1)*NF; or nullptr (plain gradient step)
  const double* __restrict__ J_d;      // [B]
  const double* __restrict__ J_sc;     // [B] or nullptr
  double* __restrict__ f_prev;         // [B]
  double* __restrict__ step;           // [B]
  double* __restrict__ accepted;       // [B] or nullptr: 1.0 / 0.0 (a row of cost_history)
  uint8_t* __restrict__ flag;          // [B]
  double bound[MTG_TAB_LD];            // |x| <= bound[k] for derivative order k (infinity: unbounded)
  double w_d, w_sc;
  int B, nb, Q, NF, per_dim;
};

__global__ void __launch_bounds__(256) descent_accept_kernel(const DescentParams p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.nb) return;
  double f = p.w_d * p.J_d[b];
  if (p.J_sc) f = fma(p.w_sc, p.J_sc[b], f);
  const bool acc = f <= p.f_prev[b];  // false for NaN
  if (acc)
    p.f_prev[b] = f;
  else
    p.step[b] *= 0.5;
  p.flag[b] = acc ? 1 : 0;
  if (p.accepted) p.accepted[b] = acc ? 1.0 : 0.0;
}

// MODE 0: take the next trial step; MODE 1: finalize (a rejected last trial falls back to x_prev);
// MODE 2: project the start point onto the bounds (the optimiser's iterates live inside them)
template <bool AOS, int MODE>
__global__ void __launch_bounds__(256) descent_step_kernel(const DescentParams p) {
  const size_t total = (size_t)p.nb * p.Q;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
    const size_t b = AOS ? g / p.Q : g % p.nb;
    const int q = (int)(AOS ? g % p.Q : g / p.nb);
    const size_t o = at<AOS>((size_t)q, (size_t)p.Q, (size_t)p.B, b);
    if (MODE == 2) {
      const double bd0 = p.bound[(q % p.per_dim) % p.NF + 1];
      p.x[o] = fmin(fmax(p.x[o], -bd0), bd0);
      continue;
    }
    const bool acc = p.flag[b] != 0;
    if (MODE == 1) {
      if (!acc) p.x[o] = p.x_prev[o];
      continue;
    }
    double xa, gr;
    if (acc) {
      xa = p.x[o];
      gr = p.w_d * p.grad_d[o];
      if (p.grad_sc) gr = fma(p.w_sc, p.grad_sc[o], gr);
      const int within = q % p.per_dim;  // (v-1)*NF + k-1
      if (p.diag) gr /= p.diag[at<AOS>((size_t)within, (size_t)p.per_dim, (size_t)p.B, b)];
      p.x_prev[o] = xa;
      p.g_prev[o] = gr;
    } else {
      xa = p.x_prev[o];
      gr = p.g_prev[o];
    }
    const int k = (q % p.per_dim) % p.NF + 1;
    double xn = xa - p.step[b] * gr;
    const double bd = p.bound[k];
    xn = fmin(fmax(xn, -bd), bd);  // setFreeEndpointDerivativeHardConstraints, NL_I:2858-2905
    p.x[o] = xn;
  }
}

__global__ void fill_kernel(double* x, double v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = v;
}

int check_nl_desc(mtg_ctx* ctx, const mtg_problem_desc* desc) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (desc->memory != MTG_MEM_DEVICE)
    return fail(ctx, MTG_ERR_UNSUPPORTED, "the non-linear objective entry points take device pointers (they run inside optimiser loops)");
  if (desc->K < 2) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "K >= 2: a single segment has no free derivatives");
  if (desc->N < 4) return fail(ctx, MTG_ERR_UNSUPPORTED, "supported N: {4,6,8,10,12}");
  return MTG_OK;
}

}  // namespace

extern "C" {

int mtg_cost_derivative_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* positions,
                              const double* end_derivatives, const double* seg_times, const double* free_constraints,
                              double* J_d, double* grad, double* diag, uint32_t* status, void* stream_) {
  int rc = check_nl_desc(ctx, desc);
  if (rc) return rc;
  if (!positions || !seg_times || !free_constraints || (!J_d && !grad && !diag))
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "positions, seg_times, free_constraints and an output are required");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  TableGuard tables(ctx, desc->N, desc->derivative_to_optimize, (cudaStream_t)stream_);
  if (tables.rc()) return tables.rc();
  CostDerivParams p = {};
  p.in.positions = positions; p.in.end_derivatives = end_derivatives; p.in.seg_times = seg_times;
  p.in.free_constraints = free_constraints; p.in.status = status;
  p.in.B = desc->B; p.in.b0 = 0; p.in.nb = desc->B; p.in.K = desc->K; p.in.derivative = desc->derivative_to_optimize;
  p.J_d = J_d; p.grad = grad; p.diag = diag;
  return desc->layout == MTG_LAYOUT_AOS ? launch_cdv_n<true>(ctx, desc->N, desc->D, p, (cudaStream_t)stream_)
                                        : launch_cdv_n<false>(ctx, desc->N, desc->D, p, (cudaStream_t)stream_);
}

int mtg_soft_constraint_gradient_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                                       const double* seg_times, int n_constraints, const int32_t* derivatives,
                                       const double* limits, double weight, double maximum_cost, double increment,
                                       int central, double* J_sc, double* grad, uint32_t* status, void* stream_) {
  int rc = check_nl_desc(ctx, desc);
  if (rc) return rc;
  if (!coeffs || !seg_times || n_constraints < 1 || n_constraints > kMaxSoft || !derivatives || !limits ||
      !(increment > 0.0) || (!J_sc && !grad))
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "coeffs, seg_times, 1..4 constraints, increment > 0 and an output are required");
  for (int c = 0; c < n_constraints; ++c)
    if (derivatives[c] < 0 || desc->N - derivatives[c] - 1 <= 0)
      return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "constraint derivative must satisfy 0 <= derivative < N - 1");
  if (desc->layout != MTG_LAYOUT_AOS)
    return fail(ctx, MTG_ERR_UNSUPPORTED, "mtg_soft_constraint_gradient_batch: AoS layout only");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream_;
  TableGuard tables(ctx, desc->N, desc->derivative_to_optimize, s);
  if (tables.rc()) return tables.rc();
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N, NF = N / 2 - 1;
  const int Q = D * (K - 1) * NF;
  // trajectories per pass: bounds the work space (perturbed coefficients: Q * 2 * 2 * D * N doubles each)
  const size_t per_traj = (size_t)Q * 2 * (2 * D * N + 2) * 8 + (size_t)n_constraints * (K + (size_t)Q * 4) * 8;
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)B, ((size_t)768 << 20) / per_traj));
  DeviceBuffer* ws = ctx->nl_scratch_for(s, 0);
  if (ws->ensure(per_traj * chunk + 4096)) return fail(ctx, MTG_ERR_CUDA, "cudaMalloc of the soft-gradient work space failed");
  char* base = (char*)ws->ptr;
  double* pc = (double*)base;
  double* pt = pc + (size_t)chunk * Q * 2 * 2 * D * N;
  double* nom0 = pt + (size_t)chunk * Q * 2 * 2;
  double* per0 = nom0 + (size_t)n_constraints * chunk * K;
  for (int off = 0; off < B; off += chunk) {
    const int nb = std::min(chunk, B - off);
    SoftCombineParams sc = {};
    // AoS: a sub-batch is a pointer offset
    const double* c_off = coeffs + (size_t)off * K * D * N;
    const double* t_off = seg_times + (size_t)off * K;
    // nominal per-segment maxima of every constraint, AoS [nb][K] in the work space
    for (int c = 0; c < n_constraints; ++c) {
      ExtremaParams e = {};
      e.K = K; e.N = N; e.D = D; e.derivative = derivatives[c]; e.dim_mask = (1 << D) - 1;
      e.coeffs = c_off; e.seg_times = t_off; e.B = nb; e.b0 = 0; e.nb = nb;
      double* nom = nom0 + (size_t)c * chunk * K;
      e.seg_max_value = nom;
      e.status = (status && c == 0) ? status + off : nullptr;
      rc = launch_extrema(ctx, true, e, s);
      if (rc) return rc;
      sc.nominal[c] = nom;
      sc.limit[c] = limits[c];
    }
    if (!grad) {  // cost only: no perturbed segments
      sc.J_sc = J_sc + off; sc.grad = nullptr; sc.weight = weight; sc.max_cost = maximum_cost; sc.increment = increment;
      sc.central = 1; sc.n_con = n_constraints; sc.B = nb; sc.b0 = 0; sc.nb = nb; sc.K = K; sc.Q = 1; sc.NF = NF;
      sc.cost_only = 1;
      soft_combine_kernel<true><<<(unsigned)((nb + 127) / 128), 128, 0, s>>>(sc);
      ++ctx->launches;
      MTG_CUDA_TRY(cudaGetLastError());
      continue;
    }
    PerturbParams pp = {};
    pp.coeffs = c_off; pp.seg_times = t_off; pp.pc = pc; pp.pt = pt; pp.increment = increment;
    pp.B = nb; pp.b0 = 0; pp.nb = nb; pp.K = K; pp.D = D; pp.N = N;
    {
      const size_t total = (size_t)nb * Q * 2 * 2 * D * N;
      const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 32);
      perturb_segments_kernel<true><<<grid, 256, 0, s>>>(pp);
      ++ctx->launches;
      MTG_CUDA_TRY(cudaGetLastError());
    }
    const int items = nb * Q * 2;
    for (int c = 0; c < n_constraints; ++c) {
      ExtremaParams e = {};
      e.K = 2; e.N = N; e.D = D; e.derivative = derivatives[c]; e.dim_mask = (1 << D) - 1;
      e.coeffs = pc; e.seg_times = pt; e.B = items; e.b0 = 0; e.nb = items;
      double* out = per0 + (size_t)c * chunk * Q * 4;
      e.seg_max_value = out;  // AoS [items][2]
      sc.perturbed[c] = out;
      rc = launch_extrema(ctx, true, e, s);
      if (rc) return rc;
    }
    sc.J_sc = J_sc ? J_sc + off : nullptr; sc.grad = grad ? grad + (size_t)off * Q : nullptr; sc.weight = weight; sc.max_cost = maximum_cost; sc.increment = increment;
    sc.central = central ? 1 : 0; sc.n_con = n_constraints; sc.B = nb; sc.b0 = 0; sc.nb = nb; sc.K = K; sc.Q = Q; sc.NF = NF;
    {
      const size_t total = (size_t)nb * Q;
      soft_combine_kernel<true><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(sc);
      ++ctx->launches;
      MTG_CUDA_TRY(cudaGetLastError());
    }
  }
  return MTG_OK;
}

int mtg_nl_descent_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* positions,
                         const double* end_derivatives, const double* seg_times, double* free_constraints,
                         int n_constraints, const int32_t* derivatives, const double* limits, double w_d, double w_sc,
                         double soft_weight, double maximum_cost, double increment, double step, int precondition,
                         int iterations, double* coeffs, double* cost_history, uint32_t* status, void* stream_) {
  int rc = check_nl_desc(ctx, desc);
  if (rc) return rc;
  if (!positions || !seg_times || !free_constraints || !coeffs || iterations < 0 || n_constraints < 0 ||
      n_constraints > kMaxSoft || (n_constraints > 0 && (!derivatives || !limits)) || !(step > 0.0))
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "positions, seg_times, free_constraints, coeffs, step > 0 and a valid constraint list are required");
  if (desc->layout != MTG_LAYOUT_AOS)
    return fail(ctx, MTG_ERR_UNSUPPORTED, "mtg_nl_descent_batch: AoS layout only");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream_;
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N, NF = N / 2 - 1;
  const int Q = D * (K - 1) * NF, per_dim = (K - 1) * NF;
  DeviceBuffer* gbuf = ctx->nl_scratch_for(s, 1);  // gradients, diagonal, optimiser state (slot 0: soft-gradient work space)
  const size_t n_d = ((size_t)4 * Q + per_dim + 4) * B;
  if (gbuf->ensure(n_d * 8 + (size_t)B + 256)) return fail(ctx, MTG_ERR_CUDA, "cudaMalloc of the optimiser state failed");
  double* g_d = (double*)gbuf->ptr;
  double* g_sc = g_d + (size_t)Q * B;
  double* x_prev = g_sc + (size_t)Q * B;
  double* g_prev = x_prev + (size_t)Q * B;
  double* dg = g_prev + (size_t)Q * B;
  double* Jd_buf = dg + (size_t)per_dim * B;
  double* Jsc_buf = Jd_buf + B;
  double* f_prev = Jsc_buf + B;
  double* step_b = f_prev + B;
  uint8_t* flag = (uint8_t*)(step_b + B);
  const bool soft = n_constraints > 0 && w_sc != 0.0;
  fill_kernel<<<(B + 255) / 256, 256, 0, s>>>(f_prev, INFINITY, B);
  fill_kernel<<<(B + 255) / 256, 256, 0, s>>>(step_b, step, B);
  ctx->launches += 2;
  DescentParams dp = {};
  dp.x = free_constraints; dp.x_prev = x_prev; dp.g_prev = g_prev; dp.grad_d = g_d; dp.grad_sc = soft ? g_sc : nullptr;
  dp.diag = precondition ? dg : nullptr; dp.f_prev = f_prev; dp.step = step_b; dp.flag = flag;
  for (int k = 0; k < MTG_TAB_LD; ++k) dp.bound[k] = INFINITY;
  for (int c = 0; c < n_constraints; ++c)
    if (derivatives[c] >= 1 && derivatives[c] <= NF) dp.bound[derivatives[c]] = std::fabs(limits[c]);
  dp.w_d = w_d; dp.w_sc = w_sc; dp.B = B; dp.nb = B; dp.Q = Q; dp.NF = NF; dp.per_dim = per_dim;
  const size_t total = (size_t)B * Q;
  const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 32);
  descent_step_kernel<true, 2><<<grid, 256, 0, s>>>(dp);  // the start point, projected onto the bounds
  ++ctx->launches;
  // rows of cost_history: [pass][3][B] = J_d, J_sc, accepted; passes 0..iterations are the trial points, pass
  // iterations + 1 is the returned (last accepted) point
  for (int it = 0; it <= iterations + 1; ++it) {
    const bool final_pass = it == iterations + 1;
    rc = mtg_set_free_constraints_batch(ctx, desc, positions, end_derivatives, seg_times, free_constraints, coeffs,
                                        nullptr, status, stream_);
    if (rc) return rc;
    double* Jd = cost_history ? cost_history + (size_t)it * 3 * B : Jd_buf;
    double* Jsc = cost_history ? Jd + B : Jsc_buf;
    rc = mtg_cost_derivative_batch(ctx, desc, positions, end_derivatives, seg_times, free_constraints, Jd,
                                   final_pass ? nullptr : g_d, (precondition && !final_pass) ? dg : nullptr, nullptr, stream_);
    if (rc) return rc;
    if (soft) {
      rc = mtg_soft_constraint_gradient_batch(ctx, desc, coeffs, seg_times, n_constraints, derivatives, limits,
                                              soft_weight, maximum_cost, increment, 1, Jsc, final_pass ? nullptr : g_sc,
                                              nullptr, stream_);
      if (rc) return rc;
    } else {
      MTG_CUDA_TRY(cudaMemsetAsync(Jsc, 0, sizeof(double) * (size_t)B, s));
    }
    if (final_pass) {
      if (cost_history) {
        fill_kernel<<<(B + 255) / 256, 256, 0, s>>>(Jd + 2 * (size_t)B, 1.0, B);
        ++ctx->launches;
      }
      break;
    }
    dp.J_d = Jd; dp.J_sc = soft ? Jsc : nullptr; dp.accepted = cost_history ? Jd + 2 * (size_t)B : nullptr;
    descent_accept_kernel<<<(B + 255) / 256, 256, 0, s>>>(dp);
    if (it == iterations)
      descent_step_kernel<true, 1><<<grid, 256, 0, s>>>(dp);
    else
      descent_step_kernel<true, 0><<<grid, 256, 0, s>>>(dp);
    ctx->launches += 2;
    MTG_CUDA_TRY(cudaGetLastError());
  }
  return MTG_OK;
}

}  // extern "C"

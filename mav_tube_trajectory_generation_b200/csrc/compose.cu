// libmtg_cuda.so — N3: batched trajectory composition and I/O in the solve's record layouts.
//
// Replaces (reference):
//   Trajectory::getVertexAtTime / getStartVertex / getGoalVertex        src/trajectory.cpp:248-262
//   Trajectory::getTrajectoryWithSingleDimension / ...AppendedDimension src/trajectory.cpp:136-182
//     (Segment::getSegmentWithSingleDimension / ...AppendedDimension    src/segment.cpp:186-222)
//   Trajectory::addTrajectories                                         src/trajectory.cpp:230-246
//   PolynomialOptimizationNonLinear::printMatlabSampledTrajectory       NL_I:2907-3003 — the sample dump
//     [t, pos, vel, acc, jerk, snap, tm] (as a tensor; writing the text file stays with the caller)
//   PolynomialOptimization::computeCost on given coefficients           LIN_I:113-130
// These are data-movement kernels (HBM-bound, no arithmetic to speak of) except the dump, which is a
// five-derivative evaluation sweep: one warp per trajectory, lane = sample, rows staged in shared memory
// and written as whole coalesced lines.
#include "host_common.h"
#include "eval.cuh"  // at<AOS>()

MTG_REGISTER_BASE()
MTG_REGISTER_TABLES()

using namespace mtg;

namespace {

// ------------------------------------------------------------------ getVertexAtTime
struct VertexAtParams {
  const double* __restrict__ coeffs;     // elem ((i*D + dim)*N + j), rec K*D*N
  const double* __restrict__ seg_times;  // elem i, rec K
  const double* __restrict__ t;          // [B]
  double* __restrict__ out;              // elem k*D + dim, rec (M+1)*D
  int32_t* __restrict__ segment_idx;     // [B] or nullptr
  uint32_t* __restrict__ status;         // [B] or nullptr
  int B, b0, nb, K, D, N, M;
};

template <bool AOS>
__global__ void __launch_bounds__(128) vertex_at_kernel(const VertexAtParams p) {
  const int local = blockIdx.x * blockDim.x + threadIdx.x;
  if (local >= p.nb) return;
  const int b = p.b0 + local;
  const size_t B = (size_t)p.B;
  const int K = p.K, D = p.D, N = p.N, M = p.M;
  const double t = p.t[b];
  // Trajectory::evaluate, trajectory.cpp:41-72
  double acc = 0.0, Ti = 0.0;
  int i = 0;
  for (i = 0; i < K; ++i) {
    Ti = p.seg_times[at<AOS>((size_t)i, (size_t)K, B, b)];
    acc += Ti;
    if (acc > t) break;
  }
  const size_t rec_o = (size_t)(M + 1) * D, rec_c = (size_t)K * D * N;
  if (t > acc || !(t == t)) {  // LOG(ERROR) + zero vector
    for (int e = 0; e < (M + 1) * D; ++e) p.out[at<AOS>((size_t)e, rec_o, B, b)] = 0.0;
    if (p.segment_idx) p.segment_idx[b] = -1;
    if (p.status) p.status[b] = 4u;  // MTG_ST_OUT_OF_RANGE
    return;
  }
  if (i >= K) i = K - 1;
  acc -= Ti;
  const double tau = t - acc;
  for (int dim = 0; dim < D; ++dim)
    for (int k = 0; k <= M; ++k) {
      double r = 0.0;  // polynomial.h:136-149
      if (k < N) {
        r = c_base.base[k * MTG_BASE_LD + (N - 1)] * p.coeffs[at<AOS>((size_t)(i * D + dim) * N + (N - 1), rec_c, B, b)];
        for (int j = N - 2; j >= k; --j)
          r = fma(r, tau, c_base.base[k * MTG_BASE_LD + j] * p.coeffs[at<AOS>((size_t)(i * D + dim) * N + j, rec_c, B, b)]);
      }
      p.out[at<AOS>((size_t)k * D + dim, rec_o, B, b)] = r;
    }
  if (p.segment_idx) p.segment_idx[b] = i;
  if (p.status) p.status[b] = 0u;
}

// ------------------------------------------------------------------ dimension selection / append
constexpr int kMaxPick = 8;
struct PickParams {
  const double* __restrict__ a;  // rec K*Da*N
  const double* __restrict__ bsrc;  // rec K*Db*N or nullptr
  double* __restrict__ out;      // rec K*n_out*N
  int pick[kMaxPick];            // < Da: dimension of a; >= Da: dimension (pick - Da) of b
  int B, b0, nb, K, N, Da, Db, n_out;
};

template <bool AOS>
__global__ void __launch_bounds__(256) pick_dimensions_kernel(const PickParams p) {
  const size_t rec_o = (size_t)p.K * p.n_out * p.N;
  const size_t total = rec_o * (size_t)p.nb;
  const size_t B = (size_t)p.B;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
    // consecutive threads -> consecutive memory in the OUTPUT layout
    const size_t local = AOS ? g / rec_o : g % p.nb;
    const size_t e = AOS ? g % rec_o : g / p.nb;
    const int j = (int)(e % p.N);
    const int q = (int)((e / p.N) % p.n_out);
    const int k = (int)(e / ((size_t)p.N * p.n_out));
    const int src = p.pick[q];
    const size_t b = p.b0 + local;
    double v;
    if (src < p.Da)
      v = p.a[at<AOS>((size_t)(k * p.Da + src) * p.N + j, (size_t)p.K * p.Da * p.N, B, b)];
    else
      v = p.bsrc[at<AOS>((size_t)(k * p.Db + (src - p.Da)) * p.N + j, (size_t)p.K * p.Db * p.N, B, b)];
    p.out[at<AOS>(e, rec_o, B, b)] = v;
  }
}

// ------------------------------------------------------------------ computeCost of given coefficients
struct CostParams {
  const double* __restrict__ coeffs;
  const double* __restrict__ seg_times;
  double* __restrict__ cost;
  uint32_t* __restrict__ status;
  int B, b0, nb, K, D, N, derivative;
};

// 0.5 sum_seg sum_dim c^T Q c (LIN_I:113-130) as a sum of squares through the triangular factor of Q(1):
// c^T Q(T) c = T^(1-2d) |Lt chat|^2 with chat_j = c_j T^j (no cancellation across terms)
template <bool AOS>
__global__ void __launch_bounds__(128) compute_cost_kernel(const CostParams p) {
  const int local = blockIdx.x * blockDim.x + threadIdx.x;
  if (local >= p.nb) return;
  const int b = p.b0 + local;
  const size_t B = (size_t)p.B;
  const int K = p.K, D = p.D, N = p.N, d = p.derivative, nq = N - d;
  const size_t rec_c = (size_t)K * D * N;
  uint32_t st = 0;
  double total = 0.0;
  for (int i = 0; i < K; ++i) {
    double T = p.seg_times[at<AOS>((size_t)i, (size_t)K, B, b)];
    if (!(T > 0.0) || !(T < 1.7e308)) {
      st |= 1u;
      T = 1.0;
    }
    double quad = 0.0;
    for (int dim = 0; dim < D; ++dim) {
      double chat[MTG_TAB_LD];
      double tp = 1.0;
      for (int j = 0; j < N; ++j) {
        chat[j] = p.coeffs[at<AOS>((size_t)(i * D + dim) * N + j, rec_c, B, b)] * tp;
        tp *= T;
      }
      for (int r = 0; r < nq; ++r) {
        double w = 0.0;
        for (int a = r; a < nq; ++a) w = fma(c_tab.Lt[r * MTG_TAB_LD + a], chat[d + a], w);
        quad = fma(w, w, quad);
      }
    }
    double s = 1.0;  // T^(1-2d)
    const int e0 = 1 - 2 * d;
    for (int q = 0; q < (e0 >= 0 ? e0 : -e0); ++q) s *= (e0 >= 0 ? T : 1.0 / T);
    total += quad * s;
  }
  p.cost[b] = 0.5 * total;
  if (p.status) p.status[b] = st;
}

// ------------------------------------------------------------------ the sample dump (NL_I:2907-3003)
struct DumpParams {
  const double* __restrict__ coeffs;     // AoS or SoA, elem ((i*D + dim)*N + j), rec K*D*N
  const double* __restrict__ seg_times;  // elem i, rec K
  double* __restrict__ rows;             // [B][max_rows][5 D + 2], trajectory-contiguous in BOTH layouts
  int32_t* __restrict__ n_rows;          // [B] or nullptr: rows that hold samples
  uint32_t* __restrict__ status;         // [B] or nullptr
  double dt;
  int B, b0, nb, K, D, N, max_rows;
};
constexpr int kDumpWarps = 4;

template <bool AOS>
__global__ void __launch_bounds__(kDumpWarps * 32) sample_dump_kernel(const DumpParams p) {
  extern __shared__ double dump_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int local = blockIdx.x * kDumpWarps + warp;
  if (local >= p.nb) return;
  const int b = p.b0 + local;
  const size_t B = (size_t)p.B;
  const int K = p.K, D = p.D, N = p.N;
  const int W = 5 * D + 2;  // row width: t, 5 derivative orders x D, tm
  double* stage = dump_smem + (size_t)warp * 32 * W;
  double* out = p.rows + (size_t)b * p.max_rows * W;
  const size_t rec_c = (size_t)K * D * N;
  uint32_t st = 0;
  int row0 = 0;          // rows written so far (the reference's j)
  double seg_start = 0.0;  // current_segment_time
  for (int i = 0; i < K; ++i) {
    const double T = p.seg_times[at<AOS>((size_t)i, (size_t)K, B, b)];
    // for (t = 0; t < T_i; t += dt): every lane replays the accumulation up to ITS sample (lane l of chunk c
    // holds t after 32 c + l additions: the same additions in the same order as the reference's loop)
    double t = 0.0;
    for (int q = 0; q < lane; ++q) t += p.dt;
    for (int c0 = 0;; c0 += 32) {
      const bool live = t < T;
      const unsigned m = __ballot_sync(0xffffffffu, live);
      if (m == 0u) break;
      const int cnt = __popc(m);  // live lanes are a prefix: t grows with the lane
      if (live) {
        double* r = stage + lane * W;
        r[0] = t + seg_start;
        for (int dim = 0; dim < D; ++dim) {
          double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
          for (int k = 0; k < 5; ++k) {
            if (k >= N) continue;
            double acc = c_base.base[k * MTG_BASE_LD + (N - 1)] * p.coeffs[at<AOS>((size_t)(i * D + dim) * N + (N - 1), rec_c, B, b)];
            for (int j = N - 2; j >= k; --j)
              acc = fma(acc, t, c_base.base[k * MTG_BASE_LD + j] * p.coeffs[at<AOS>((size_t)(i * D + dim) * N + j, rec_c, B, b)]);
            v[k] = acc;
          }
          for (int k = 0; k < 5; ++k) r[1 + k * D + dim] = v[k];
        }
        r[W - 1] = 0.0;
      }
      __syncwarp();
      // coalesced copy of the chunk's rows (rows beyond max_rows are dropped: the reference's `if (j < rows)`)
      const int fit = max(0, min(cnt, p.max_rows - row0));
      for (int e = lane; e < fit * W; e += 32) out[(size_t)row0 * W + e] = stage[e];
      if (fit < cnt) st |= 8u;  // MTG_ST_TRUNCATED
      row0 += fit;
      __syncwarp();
      for (int q = 0; q < 32; ++q) t += p.dt;
    }
    seg_start += T;
  }
  __syncwarp();
  // rows without samples are zero (output.setZero(), NL_I:2943)
  for (size_t e = (size_t)row0 * W + lane; e < (size_t)p.max_rows * W; e += 32) out[e] = 0.0;
  __syncwarp();
  // output(i, 1 + 5 D) = end time of segment i, in ROW i (NL_I:2994-2995)
  if (lane == 0) {
    double acc = 0.0;
    for (int i = 0; i < K; ++i) {
      acc += p.seg_times[at<AOS>((size_t)i, (size_t)K, B, b)];
      if (i < p.max_rows) out[(size_t)i * W + (W - 1)] = acc;
    }
    if (p.n_rows) p.n_rows[b] = row0;
    if (p.status) p.status[b] = st;
  }
}

}  // namespace

extern "C" {

int mtg_vertex_at_time_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                             const double* seg_times, const double* t, int max_derivative_order, double* out,
                             int32_t* segment_idx, uint32_t* status, void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (!coeffs || !seg_times || !t || !out || max_derivative_order < 0)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "coeffs, seg_times, t, out and max_derivative_order >= 0 are required");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N, M = max_derivative_order;
  VertexAtParams p = {};
  p.K = K; p.D = D; p.N = N; p.M = M;
  auto launch = [&](cudaStream_t st) {
    const int grid = (p.nb + 127) / 128;
    if (aos) vertex_at_kernel<true><<<grid, 128, 0, st>>>(p); else vertex_at_kernel<false><<<grid, 128, 0, st>>>(p);
    ++ctx->launches;
    MTG_CUDA_TRY(cudaGetLastError());
    return (int)MTG_OK;
  };
  if (desc->memory == MTG_MEM_DEVICE) {
    p.coeffs = coeffs; p.seg_times = seg_times; p.t = t; p.out = out; p.segment_idx = segment_idx; p.status = status;
    p.B = B; p.b0 = 0; p.nb = B;
    return launch(stream);
  }
  std::vector<HostTensor> ts = {
      {coeffs, (size_t)K * D * N, 8, true, false, nullptr}, {seg_times, (size_t)K, 8, true, false, nullptr},
      {t, 1, 8, true, true, nullptr}, {out, (size_t)(M + 1) * D, 8, false, false, nullptr},
      {segment_idx, 1, 4, false, true, nullptr}, {status, 1, 4, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    p.coeffs = (const double*)ts[0].dev; p.seg_times = (const double*)ts[1].dev; p.t = (const double*)ts[2].dev;
    p.out = (double*)ts[3].dev; p.segment_idx = (int32_t*)ts[4].dev; p.status = (uint32_t*)ts[5].dev;
    p.B = C; p.b0 = 0; p.nb = nb;
    return launch(st);
  });
}

int mtg_pick_dimensions_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs_a, int D_b,
                              const double* coeffs_b, int n_out, const int32_t* pick, double* coeffs_out,
                              void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (!coeffs_a || !coeffs_out || !pick || n_out < 1 || n_out > kMaxPick || D_b < 0 || (D_b > 0 && !coeffs_b))
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "coeffs_a, pick[1..8], coeffs_out (and coeffs_b when D_b > 0) are required");
  for (int q = 0; q < n_out; ++q)
    if (pick[q] < 0 || pick[q] >= desc->D + D_b)  // CHECK_LT(dimension, D_), trajectory.cpp:137
      return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "pick entry outside [0, D_a + D_b)");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  const int B = desc->B, K = desc->K, N = desc->N;
  PickParams p = {};
  p.K = K; p.N = N; p.Da = desc->D; p.Db = D_b; p.n_out = n_out;
  for (int q = 0; q < n_out; ++q) p.pick[q] = pick[q];
  auto launch = [&](cudaStream_t st) {
    const size_t total = (size_t)K * n_out * N * p.nb;
    const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 32);
    if (aos) pick_dimensions_kernel<true><<<grid, 256, 0, st>>>(p); else pick_dimensions_kernel<false><<<grid, 256, 0, st>>>(p);
    ++ctx->launches;
    MTG_CUDA_TRY(cudaGetLastError());
    return (int)MTG_OK;
  };
  if (desc->memory == MTG_MEM_DEVICE) {
    p.a = coeffs_a; p.bsrc = coeffs_b; p.out = coeffs_out; p.B = B; p.b0 = 0; p.nb = B;
    return launch(stream);
  }
  std::vector<HostTensor> ts = {{coeffs_a, (size_t)K * p.Da * N, 8, true, false, nullptr},
                                {D_b ? coeffs_b : nullptr, (size_t)K * std::max(D_b, 1) * N, 8, true, false, nullptr},
                                {coeffs_out, (size_t)K * n_out * N, 8, false, false, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    p.a = (const double*)ts[0].dev; p.bsrc = (const double*)ts[1].dev; p.out = (double*)ts[2].dev;
    p.B = C; p.b0 = 0; p.nb = nb;
    return launch(st);
  });
}

// addTrajectories: the segments of trajectory b of every input, one after the other. Pure copies: strided
// (2-D) DMA transfers, no kernel — AoS: per-record concatenation; SoA: the element-major arrays stacked.
int mtg_concat_segments_batch(mtg_ctx* ctx, int B, int D, int N, int memory, int layout, int n_inputs,
                              const int32_t* K_in, const double* const* coeffs_in, const double* const* times_in,
                              double* coeffs_out, double* times_out, void* stream_) {
  if (!ctx) return MTG_ERR_INVALID_ARGUMENT;
  if (B < 0 || D < 1 || N < 1 || n_inputs < 1 || !K_in || !coeffs_in || !times_in || !coeffs_out || !times_out)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "inputs and outputs are required");
  if ((memory != MTG_MEM_DEVICE && memory != MTG_MEM_HOST) || (layout != MTG_LAYOUT_SOA && layout != MTG_LAYOUT_AOS))
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "bad memory / layout");
  if (B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream_;
  size_t K_total = 0;
  for (int q = 0; q < n_inputs; ++q) {
    if (K_in[q] < 0 || !coeffs_in[q] || !times_in[q]) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "bad input trajectory set");
    K_total += (size_t)K_in[q];
  }
  const cudaMemcpyKind kind = cudaMemcpyDefault;
  size_t k0 = 0;
  for (int q = 0; q < n_inputs; ++q) {
    const size_t Kq = (size_t)K_in[q];
    if (Kq == 0) continue;
    if (layout == MTG_LAYOUT_AOS) {
      const size_t rc_in = Kq * D * N * 8, rc_out = K_total * D * N * 8;
      MTG_CUDA_TRY(cudaMemcpy2DAsync((char*)coeffs_out + k0 * D * N * 8, rc_out, coeffs_in[q], rc_in, rc_in, (size_t)B, kind, s));
      MTG_CUDA_TRY(cudaMemcpy2DAsync((char*)times_out + k0 * 8, K_total * 8, times_in[q], Kq * 8, Kq * 8, (size_t)B, kind, s));
    } else {
      MTG_CUDA_TRY(cudaMemcpyAsync(coeffs_out + k0 * D * N * (size_t)B, coeffs_in[q], Kq * D * N * (size_t)B * 8, kind, s));
      MTG_CUDA_TRY(cudaMemcpyAsync(times_out + k0 * (size_t)B, times_in[q], Kq * (size_t)B * 8, kind, s));
    }
    k0 += Kq;
  }
  if (memory == MTG_MEM_HOST) MTG_CUDA_TRY(cudaStreamSynchronize(s));
  return MTG_OK;
}

int mtg_compute_cost_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs, const double* seg_times,
                           double* cost, uint32_t* status, void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (!coeffs || !seg_times || !cost) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "coeffs, seg_times and cost are required");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  TableGuard tables(ctx, desc->N, desc->derivative_to_optimize, (cudaStream_t)stream_);
  if (tables.rc()) return tables.rc();
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N;
  CostParams p = {};
  p.K = K; p.D = D; p.N = N; p.derivative = desc->derivative_to_optimize;
  auto launch = [&](cudaStream_t st) {
    const int grid = (p.nb + 127) / 128;
    if (aos) compute_cost_kernel<true><<<grid, 128, 0, st>>>(p); else compute_cost_kernel<false><<<grid, 128, 0, st>>>(p);
    ++ctx->launches;
    MTG_CUDA_TRY(cudaGetLastError());
    return (int)MTG_OK;
  };
  if (desc->memory == MTG_MEM_DEVICE) {
    p.coeffs = coeffs; p.seg_times = seg_times; p.cost = cost; p.status = status; p.B = B; p.b0 = 0; p.nb = B;
    return launch(stream);
  }
  std::vector<HostTensor> ts = {{coeffs, (size_t)K * D * N, 8, true, false, nullptr},
                                {seg_times, (size_t)K, 8, true, false, nullptr},
                                {cost, 1, 8, false, true, nullptr}, {status, 1, 4, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    p.coeffs = (const double*)ts[0].dev; p.seg_times = (const double*)ts[1].dev; p.cost = (double*)ts[2].dev;
    p.status = (uint32_t*)ts[3].dev; p.B = C; p.b0 = 0; p.nb = nb;
    return launch(st);
  });
}

int mtg_sample_dump_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs, const double* seg_times,
                          double dt, int max_rows, double* rows, int32_t* n_rows, uint32_t* status, void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (!coeffs || !seg_times || !rows || !(dt > 0.0) || max_rows < 1)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "coeffs, seg_times, rows, dt > 0 and max_rows >= 1 are required");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N;
  const int W = 5 * D + 2;
  DumpParams p = {};
  p.K = K; p.D = D; p.N = N; p.dt = dt; p.max_rows = max_rows;
  const size_t smem = (size_t)kDumpWarps * 32 * W * sizeof(double);
  auto launch = [&](cudaStream_t st) {
    const int grid = (p.nb + kDumpWarps - 1) / kDumpWarps;
    if (aos) sample_dump_kernel<true><<<grid, kDumpWarps * 32, smem, st>>>(p);
    else sample_dump_kernel<false><<<grid, kDumpWarps * 32, smem, st>>>(p);
    ++ctx->launches;
    MTG_CUDA_TRY(cudaGetLastError());
    return (int)MTG_OK;
  };
  if (desc->memory == MTG_MEM_DEVICE) {
    p.coeffs = coeffs; p.seg_times = seg_times; p.rows = rows; p.n_rows = n_rows; p.status = status;
    p.B = B; p.b0 = 0; p.nb = B;
    return launch(stream);
  }
  // rows are trajectory-contiguous in both layouts: a [B] "vector" of records of max_rows * W doubles
  std::vector<HostTensor> ts = {{coeffs, (size_t)K * D * N, 8, true, false, nullptr},
                                {seg_times, (size_t)K, 8, true, false, nullptr},
                                {rows, (size_t)max_rows * W, 8, false, true, nullptr},
                                {n_rows, 1, 4, false, true, nullptr}, {status, 1, 4, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    p.coeffs = (const double*)ts[0].dev; p.seg_times = (const double*)ts[1].dev; p.rows = (double*)ts[2].dev;
    p.n_rows = (int32_t*)ts[3].dev; p.status = (uint32_t*)ts[4].dev; p.B = C; p.b0 = 0; p.nb = nb;
    return launch(st);
  });
}

}  // extern "C"

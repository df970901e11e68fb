// libmtg_cuda.so — E6 / R1: mtg_extrema_batch (analytic extrema of the derivative magnitude).
#include "host_common.h"
#include "extrema.cuh"

MTG_REGISTER_BASE()

using namespace mtg;

namespace {
constexpr int kExtremaChunk = 1 << 18;  // trajectories per launch pair (bounds the scratch: 36 B x K each)

}  // namespace

namespace mtg {
// also used by nl_objective.cu (soft-constraint gradient)
int launch_extrema(mtg_ctx* ctx, bool aos, const ExtremaParams& p_in, cudaStream_t s) {
  ExtremaParams p = p_in;
  const ExtremaPlan pl = extrema_plan(p.N, p.D, p.derivative, p.dim_mask, p.raw);
  if (pl.len > kMaxG) return fail(ctx, MTG_ERR_UNSUPPORTED, "polynomial too long for the root kernel (22 coefficients)");
  if (pl.cta_bytes > ctx->smem_optin) return fail(ctx, MTG_ERR_UNSUPPORTED, "extrema: shared memory plan exceeds the device limit");
  void (*kern)(const ExtremaParams) = aos ? extrema_warp_kernel<true> : extrema_warp_kernel<false>;
  // the reference's default problem (PolynomialOptimization<10>, 3 dimensions; velocity ... snap): compile-time sizes
  if (!p.raw && !p.t_lo && p.N == 10 && p.D == 3 && p.dim_mask == 7) {
    switch (p.derivative) {
      case 1: kern = aos ? extrema_warp_kernel<true, 10, 3, 1> : extrema_warp_kernel<false, 10, 3, 1>; break;
      case 2: kern = aos ? extrema_warp_kernel<true, 10, 3, 2> : extrema_warp_kernel<false, 10, 3, 2>; break;
      case 3: kern = aos ? extrema_warp_kernel<true, 10, 3, 3> : extrema_warp_kernel<false, 10, 3, 3>; break;
      case 4: kern = aos ? extrema_warp_kernel<true, 10, 3, 4> : extrema_warp_kernel<false, 10, 3, 4>; break;
      default: break;
    }
  }
  MTG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.cta_bytes));
  const bool reduce = !p.raw && (p.min_value || p.min_time || p.min_seg || p.max_value || p.max_time || p.max_seg ||
                                 p.seg_max_value || p.seg_max_time || p.status || p.soft_cost || p.soft_violation);
  const int chunk = std::min(p_in.nb, kExtremaChunk);
  if (reduce) {
    DeviceBuffer* scratch = ctx->scratch_for(s);
    const size_t out_bytes = align256((size_t)chunk * p.K * 4 * sizeof(double));
    if (scratch->ensure(out_bytes + (size_t)chunk * p.K * sizeof(uint32_t)))
      return fail(ctx, MTG_ERR_CUDA, "cudaMalloc of the extrema scratch failed");
    p.seg_out = (double*)scratch->ptr;
    p.seg_status = (uint32_t*)((char*)scratch->ptr + out_bytes);
  } else {
    p.seg_out = nullptr;
    p.seg_status = nullptr;
  }
  for (int off = 0; off < p_in.nb; off += chunk) {
    p.b0 = p_in.b0 + off;
    p.nb = std::min(chunk, p_in.nb - off);
    const long long n_groups = aos ? ((long long)p.nb * p.K + kExG - 1) / kExG
                                   : (long long)((p.nb + kExG - 1) / kExG) * p.K;
    const unsigned grid = (unsigned)((n_groups + kExWarps - 1) / kExWarps);
    kern<<<grid, kExWarps * 32, pl.cta_bytes, s>>>(p);
    ++ctx->launches;
    MTG_CUDA_TRY(cudaGetLastError());
    if (!reduce) continue;
    if (aos)
      extrema_reduce_kernel<true><<<(p.nb + 255) / 256, 256, 0, s>>>(p);
    else
      extrema_reduce_kernel<false><<<(p.nb + 255) / 256, 256, 0, s>>>(p);
    ++ctx->launches;
    MTG_CUDA_TRY(cudaGetLastError());
  }
  return MTG_OK;
}
}  // namespace mtg

namespace {
int validate_extrema(mtg_ctx* ctx, const mtg_problem_desc* desc, int derivative) {
  // LIN_I:400-401 CHECK(N - derivative - 1 > 0)
  if (derivative < 0 || desc->N - derivative - 1 <= 0)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "derivative must satisfy 0 <= derivative < N - 1");
  return MTG_OK;
}
}  // namespace

extern "C" int mtg_extrema_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                                 const double* seg_times, int derivative, double* min_value, double* min_time,
                                 int32_t* min_seg, double* max_value, double* max_time, int32_t* max_seg,
                                 double* seg_max_value, double* seg_max_time, uint32_t* status, void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (!coeffs || !seg_times) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "coeffs and seg_times are required");
  rc = validate_extrema(ctx, desc, derivative);
  if (rc) return rc;
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N;
  ExtremaParams p = {};
  p.K = K; p.N = N; p.D = D; p.derivative = derivative;
  p.dim_mask = (1 << D) - 1;
  if (desc->memory == MTG_MEM_DEVICE) {
    p.coeffs = coeffs; p.seg_times = seg_times;
    p.min_value = min_value; p.min_time = min_time; p.min_seg = min_seg;
    p.max_value = max_value; p.max_time = max_time; p.max_seg = max_seg;
    p.seg_max_value = seg_max_value; p.seg_max_time = seg_max_time; p.status = status;
    p.B = B; p.b0 = 0; p.nb = B;
    return launch_extrema(ctx, aos, p, stream);
  }
  std::vector<HostTensor> ts = {
      {coeffs, (size_t)K * D * N, 8, true, false, nullptr}, {seg_times, (size_t)K, 8, true, false, nullptr},
      {min_value, 1, 8, false, true, nullptr}, {min_time, 1, 8, false, true, nullptr},
      {min_seg, 1, 4, false, true, nullptr},   {max_value, 1, 8, false, true, nullptr},
      {max_time, 1, 8, false, true, nullptr},  {max_seg, 1, 4, false, true, nullptr},
      {seg_max_value, (size_t)K, 8, false, false, nullptr}, {seg_max_time, (size_t)K, 8, false, false, nullptr},
      {status, 1, 4, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    p.coeffs = (const double*)ts[0].dev; p.seg_times = (const double*)ts[1].dev;
    p.min_value = (double*)ts[2].dev; p.min_time = (double*)ts[3].dev; p.min_seg = (int32_t*)ts[4].dev;
    p.max_value = (double*)ts[5].dev; p.max_time = (double*)ts[6].dev; p.max_seg = (int32_t*)ts[7].dev;
    p.seg_max_value = (double*)ts[8].dev; p.seg_max_time = (double*)ts[9].dev; p.status = (uint32_t*)ts[10].dev;
    p.B = C; p.b0 = 0; p.nb = nb;
    return launch_extrema(ctx, aos, p, st);
  });
}

// E6 candidate lists: Segment::computeMinMaxMagnitudeCandidateTimes / ...Candidates [src/segment.cpp:82-158]
extern "C" int mtg_extrema_candidates_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                                            const double* seg_times, const double* t_start, const double* t_end,
                                            int derivative, int dim_mask, int max_candidates, double* cand_time,
                                            double* cand_value, int32_t* n_candidates, uint32_t* status,
                                            void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (!coeffs || !seg_times) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "coeffs and seg_times are required");
  rc = validate_extrema(ctx, desc, derivative);
  if (rc) return rc;
  if (max_candidates < 2 || (!cand_time && !cand_value && !n_candidates))
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "max_candidates >= 2 and at least one output are required");
  const int D = desc->D;
  if (dim_mask == 0) dim_mask = (1 << D) - 1;
  if (dim_mask < 0 || dim_mask >= (1 << D))  // segment.cpp:97-102: dimensions out of bounds
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "dim_mask selects a dimension outside [0, D)");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  const int B = desc->B, K = desc->K, N = desc->N;
  ExtremaParams p = {};
  p.K = K; p.N = N; p.D = D; p.derivative = derivative;
  p.dim_mask = dim_mask; p.max_cand = max_candidates;
  if (desc->memory == MTG_MEM_DEVICE) {
    p.coeffs = coeffs; p.seg_times = seg_times; p.t_lo = t_start; p.t_hi = t_end;
    p.cand_time = cand_time; p.cand_value = cand_value; p.n_cand = n_candidates; p.status = status;
    p.B = B; p.b0 = 0; p.nb = B;
    return launch_extrema(ctx, aos, p, stream);
  }
  const size_t MC = (size_t)max_candidates;
  std::vector<HostTensor> ts = {
      {coeffs, (size_t)K * D * N, 8, true, false, nullptr}, {seg_times, (size_t)K, 8, true, false, nullptr},
      {t_start, (size_t)K, 8, true, false, nullptr},        {t_end, (size_t)K, 8, true, false, nullptr},
      {cand_time, (size_t)K * MC, 8, false, false, nullptr}, {cand_value, (size_t)K * MC, 8, false, false, nullptr},
      {n_candidates, (size_t)K, 4, false, false, nullptr},  {status, 1, 4, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    p.coeffs = (const double*)ts[0].dev; p.seg_times = (const double*)ts[1].dev;
    p.t_lo = (const double*)ts[2].dev; p.t_hi = (const double*)ts[3].dev;
    p.cand_time = (double*)ts[4].dev; p.cand_value = (double*)ts[5].dev; p.n_cand = (int32_t*)ts[6].dev;
    p.status = (uint32_t*)ts[7].dev;
    p.B = C; p.b0 = 0; p.nb = nb;
    return launch_extrema(ctx, aos, p, st);
  });
}

// R1 root-list interface: the real roots of B polynomials inside [t_lo, t_hi]
// (findRootsJenkinsTraub + the selection of polynomial.cpp:46-60) [src/rpoly/rpoly_ak1.cpp:70-117]
extern "C" int mtg_poly_real_roots_batch(mtg_ctx* ctx, int B, int n_coeffs, int memory, int layout,
                                         const double* coeffs, const double* t_lo, const double* t_hi, int max_roots,
                                         double* roots, int32_t* n_roots, uint32_t* status, void* stream_) {
  if (!ctx) return MTG_ERR_INVALID_ARGUMENT;
  if (B < 0 || n_coeffs < 1 || n_coeffs > kMaxG)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "1 <= n_coeffs <= 22 (Polynomial::kMaxConvolutionSize)");
  if (!coeffs || !t_lo || !t_hi || max_roots < 1 || (!roots && !n_roots))
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "coeffs, t_lo, t_hi, max_roots >= 1 and an output are required");
  if ((memory != MTG_MEM_DEVICE && memory != MTG_MEM_HOST) || (layout != MTG_LAYOUT_SOA && layout != MTG_LAYOUT_AOS))
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "bad memory / layout");
  if (B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = layout == MTG_LAYOUT_AOS;
  ExtremaParams p = {};
  p.K = 1; p.N = n_coeffs; p.D = 1; p.derivative = 0; p.dim_mask = 1; p.raw = 1; p.max_cand = max_roots;
  if (memory == MTG_MEM_DEVICE) {
    p.coeffs = coeffs; p.t_lo = t_lo; p.t_hi = t_hi;
    p.cand_time = roots; p.n_cand = n_roots; p.status = status;
    p.B = B; p.b0 = 0; p.nb = B;
    return launch_extrema(ctx, aos, p, stream);
  }
  std::vector<HostTensor> ts = {
      {coeffs, (size_t)n_coeffs, 8, true, false, nullptr}, {t_lo, 1, 8, true, true, nullptr},
      {t_hi, 1, 8, true, true, nullptr},                   {roots, (size_t)max_roots, 8, false, false, nullptr},
      {n_roots, 1, 4, false, true, nullptr},               {status, 1, 4, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    p.coeffs = (const double*)ts[0].dev; p.t_lo = (const double*)ts[1].dev; p.t_hi = (const double*)ts[2].dev;
    p.cand_time = (double*)ts[3].dev; p.n_cand = (int32_t*)ts[4].dev; p.status = (uint32_t*)ts[5].dev;
    p.B = C; p.b0 = 0; p.nb = nb;
    return launch_extrema(ctx, aos, p, st);
  });
}

// E5 soft form: evaluateMaximumMagnitudeAsSoftConstraint [NL_I:2735-2766] over evaluateMaximumMagnitudeConstraint
// [NL_I:2686-2733]: sum over the constraints (derivative_c, limit_c) of min(maximum_cost, exp((max_c - limit_c) / limit_c * weight))
extern "C" int mtg_soft_constraint_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                                         const double* seg_times, int n_constraints, const int32_t* derivatives,
                                         const double* limits, double weight, double maximum_cost, double* cost,
                                         double* violations, uint32_t* status, void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (!coeffs || !seg_times || !cost || n_constraints < 0 || (n_constraints > 0 && (!derivatives || !limits)))
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "coeffs, seg_times, cost and the constraint list are required");
  if (desc->memory != MTG_MEM_DEVICE)
    return fail(ctx, MTG_ERR_UNSUPPORTED, "mtg_soft_constraint_batch takes device pointers (it runs inside optimiser loops)");
  for (int c = 0; c < n_constraints; ++c) {
    rc = validate_extrema(ctx, desc, derivatives[c]);
    if (rc) return rc;
  }
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  if (n_constraints == 0) {
    MTG_CUDA_TRY(cudaMemsetAsync(cost, 0, sizeof(double) * (size_t)desc->B, stream));
    if (status) MTG_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(uint32_t) * (size_t)desc->B, stream));
    return MTG_OK;
  }
  for (int c = 0; c < n_constraints; ++c) {
    ExtremaParams p = {};
    p.K = desc->K; p.N = desc->N; p.D = desc->D; p.derivative = derivatives[c];
    p.dim_mask = (1 << desc->D) - 1;
    p.coeffs = coeffs; p.seg_times = seg_times;
    p.B = desc->B; p.b0 = 0; p.nb = desc->B;
    p.soft_cost = cost;
    p.soft_violation = violations ? violations + (size_t)c * desc->B : nullptr;
    p.soft_limit = limits[c]; p.soft_weight = weight; p.soft_max = maximum_cost;
    p.soft_accumulate = c > 0;
    p.status = status;
    rc = launch_extrema(ctx, aos, p, stream);
    if (rc) return rc;
  }
  return MTG_OK;
}

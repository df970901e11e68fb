// libmtg_cuda.so — E1..E5, T1: sampled evaluation and feasibility sweep entry points.
#include "host_common.h"
#include "eval.cuh"

MTG_REGISTER_BASE()

using namespace mtg;

namespace {

template <int NT, bool AOS>
int launch_eval_range_d(mtg_ctx* ctx, int D, const mtg::EvalParams& p, cudaStream_t s) {
  const int block = 128, grid = (p.nb + block - 1) / block;
  if (grid == 0) return MTG_OK;
  switch (D) {
    case 1: mtg::eval_range_kernel<NT, 1, AOS><<<grid, block, 0, s>>>(p); break;
    case 2: mtg::eval_range_kernel<NT, 2, AOS><<<grid, block, 0, s>>>(p); break;
    case 3: mtg::eval_range_kernel<NT, 3, AOS><<<grid, block, 0, s>>>(p); break;
    case 4: mtg::eval_range_kernel<NT, 4, AOS><<<grid, block, 0, s>>>(p); break;
    default: return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "D must be 1..4");
  }
  ++ctx->launches;
  MTG_CUDA_TRY(cudaGetLastError());
  return MTG_OK;
}
int launch_eval_range(mtg_ctx* ctx, int D, bool aos, const mtg::EvalParams& p, cudaStream_t s) {
  // trajectory-contiguous outputs: the time-major sweep (eval_tm.cuh); MTG_EVAL_LEGACY=1 keeps
  // the one-thread-per-trajectory kernel for A/B measurements
  if (aos && mtg::eval_tm_supported(p) && !std::getenv("MTG_EVAL_LEGACY")) return mtg::launch_eval_tm(ctx, D, false, p, s);
  if (p.N == 10) return aos ? launch_eval_range_d<10, true>(ctx, D, p, s) : launch_eval_range_d<10, false>(ctx, D, p, s);
  return aos ? launch_eval_range_d<12, true>(ctx, D, p, s) : launch_eval_range_d<12, false>(ctx, D, p, s);
}

template <int NT, bool AOS>
int launch_feasibility_d(mtg_ctx* ctx, int D, const mtg::EvalParams& p, cudaStream_t s) {
  const int block = 128, grid = (p.nb + block - 1) / block;
  if (grid == 0) return MTG_OK;
  switch (D) {
    case 1: mtg::feasibility_kernel<NT, 1, AOS><<<grid, block, 0, s>>>(p); break;
    case 2: mtg::feasibility_kernel<NT, 2, AOS><<<grid, block, 0, s>>>(p); break;
    case 3: mtg::feasibility_kernel<NT, 3, AOS><<<grid, block, 0, s>>>(p); break;
    case 4: mtg::feasibility_kernel<NT, 4, AOS><<<grid, block, 0, s>>>(p); break;
    default: return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "D must be 1..4");
  }
  ++ctx->launches;
  MTG_CUDA_TRY(cudaGetLastError());
  return MTG_OK;
}
int launch_feasibility(mtg_ctx* ctx, int D, bool aos, const mtg::EvalParams& p, cudaStream_t s) {
  if (aos && mtg::eval_tm_supported(p) && !std::getenv("MTG_EVAL_LEGACY")) return mtg::launch_eval_tm(ctx, D, true, p, s);
  if (p.N == 10) return aos ? launch_feasibility_d<10, true>(ctx, D, p, s) : launch_feasibility_d<10, false>(ctx, D, p, s);
  return aos ? launch_feasibility_d<12, true>(ctx, D, p, s) : launch_feasibility_d<12, false>(ctx, D, p, s);
}

int validate_eval(mtg_ctx* ctx, const mtg_problem_desc* desc, int derivative, int max_samples) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (derivative < 0) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "derivative must be >= 0");
  if (max_samples < 0) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "max_samples must be >= 0");
  return MTG_OK;
}

}  // namespace

extern "C" {

/* ------------------------------------------------------------ evaluation */
int mtg_max_time_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* seg_times, double* max_time,
                       void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (!seg_times || !max_time) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "seg_times and max_time are required");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  const int K = desc->K;
  auto launch = [&](const double* t, double* out, int Bld, int nb, cudaStream_t st) {
    const int block = 256, grid = (nb + block - 1) / block;
    if (aos)
      mtg::max_time_kernel<true><<<grid, block, 0, st>>>(t, out, Bld, 0, nb, K);
    else
      mtg::max_time_kernel<false><<<grid, block, 0, st>>>(t, out, Bld, 0, nb, K);
    ++ctx->launches;
    return cudaGetLastError() == cudaSuccess ? MTG_OK : fail(ctx, MTG_ERR_CUDA, "max_time_kernel launch failed");
  };
  if (desc->memory == MTG_MEM_DEVICE) return launch(seg_times, max_time, desc->B, desc->B, stream);
  std::vector<HostTensor> ts = {{seg_times, (size_t)K, 8, true, false, nullptr}, {max_time, 1, 8, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)desc->B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    return launch((const double*)ts[0].dev, (double*)ts[1].dev, C, nb, st);
  });
}

int mtg_eval_range_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                         const double* seg_times, const double* t_start, const double* t_end,
                         const double* dt, int derivative, int max_samples, double* samples,
                         double* sampling_times, int32_t* segment_idx, int32_t* n_samples,
                         uint32_t* status, void* stream_) {
  int rc = validate_eval(ctx, desc, derivative, max_samples);
  if (rc) return rc;
  if (!coeffs || !seg_times || !t_start || !t_end || !dt)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "coeffs, seg_times, t_start, t_end and dt are required");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N;
  mtg::EvalParams p = {};
  p.K = K;
  p.N = N;
  p.derivative = derivative;
  p.max_samples = max_samples;
  if (desc->memory == MTG_MEM_DEVICE) {
    p.coeffs = coeffs; p.seg_times = seg_times; p.t_start = t_start; p.t_end = t_end; p.dt = dt;
    p.samples = samples; p.sampling_times = sampling_times; p.segment_idx = segment_idx;
    p.n_samples = n_samples; p.status = status;
    p.B = B; p.b0 = 0; p.nb = B;
    p.vec_ok = ((uintptr_t)coeffs % 16 == 0) ? 1 : 0;
    return launch_eval_range(ctx, D, aos, p, stream);
  }
  const size_t S = (size_t)max_samples;
  std::vector<HostTensor> ts = {
      {coeffs, (size_t)K * D * N, 8, true, false, nullptr}, {seg_times, (size_t)K, 8, true, false, nullptr},
      {t_start, 1, 8, true, true, nullptr},  {t_end, 1, 8, true, true, nullptr},
      {dt, 1, 8, true, true, nullptr},       {S ? samples : nullptr, S * D, 8, false, false, nullptr},
      {S ? sampling_times : nullptr, S, 8, false, false, nullptr},
      {S ? segment_idx : nullptr, S, 4, false, false, nullptr},
      {n_samples, 1, 4, false, true, nullptr}, {status, 1, 4, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    p.coeffs = (const double*)ts[0].dev; p.seg_times = (const double*)ts[1].dev;
    p.t_start = (const double*)ts[2].dev; p.t_end = (const double*)ts[3].dev; p.dt = (const double*)ts[4].dev;
    p.samples = (double*)ts[5].dev; p.sampling_times = (double*)ts[6].dev; p.segment_idx = (int32_t*)ts[7].dev;
    p.n_samples = (int32_t*)ts[8].dev; p.status = (uint32_t*)ts[9].dev;
    p.B = C; p.b0 = 0; p.nb = nb; p.vec_ok = 1;
    return launch_eval_range(ctx, D, aos, p, st);
  });
}

int mtg_eval_at_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                      const double* seg_times, const double* t, int M, int derivative, double* out,
                      int32_t* segment_idx, uint32_t* status, void* stream_) {
  int rc = validate_eval(ctx, desc, derivative, M);
  if (rc) return rc;
  if (!coeffs || !seg_times || !t || !out)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "coeffs, seg_times, t and out are required");
  if (desc->B == 0 || M == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N;
  mtg::EvalAtParams p = {};
  p.K = K; p.N = N; p.D = D; p.M = M; p.derivative = derivative;
  auto launch = [&](cudaStream_t st) {
    if (p.status) {
      cudaError_t e = cudaMemsetAsync(p.status, 0, sizeof(uint32_t) * p.nb, st);
      if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMemsetAsync(status)");
    }
    const long long total = (long long)p.nb * M;
    const int block = 256;
    const long long grid = (total + block - 1) / block;
    if (aos)
      mtg::eval_at_kernel<true><<<(unsigned)grid, block, 0, st>>>(p);
    else
      mtg::eval_at_kernel<false><<<(unsigned)grid, block, 0, st>>>(p);
    ++ctx->launches;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? MTG_OK : cuda_fail(ctx, e, "eval_at_kernel");
  };
  if (desc->memory == MTG_MEM_DEVICE) {
    p.coeffs = coeffs; p.seg_times = seg_times; p.t = t; p.out = out; p.segment_idx = segment_idx;
    p.status = status; p.B = B; p.b0 = 0; p.nb = B;
    return launch(stream);
  }
  std::vector<HostTensor> ts = {
      {coeffs, (size_t)K * D * N, 8, true, false, nullptr}, {seg_times, (size_t)K, 8, true, false, nullptr},
      {t, (size_t)M, 8, true, false, nullptr},              {out, (size_t)M * D, 8, false, false, nullptr},
      {segment_idx, (size_t)M, 4, false, false, nullptr},   {status, 1, 4, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    p.coeffs = (const double*)ts[0].dev; p.seg_times = (const double*)ts[1].dev; p.t = (const double*)ts[2].dev;
    p.out = (double*)ts[3].dev; p.segment_idx = (int32_t*)ts[4].dev; p.status = (uint32_t*)ts[5].dev;
    p.B = C; p.b0 = 0; p.nb = nb;
    return launch(st);
  });
}

int mtg_feasibility_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                          const double* seg_times, const double* positions, const double* radii,
                          double v_max, double a_max, const double* t_start, const double* t_end,
                          const double* dt, int max_samples, double* samples, uint8_t* flags,
                          double* max_v, double* max_a, uint8_t* feasible, int32_t* n_samples,
                          uint32_t* status, void* stream_) {
  int rc = validate_eval(ctx, desc, 0, max_samples);
  if (rc) return rc;
  if (!coeffs || !seg_times || !t_start || !t_end || !dt)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "coeffs, seg_times, t_start, t_end and dt are required");
  if (radii && (desc->D != 3 || !positions))
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "the tube check needs D == 3 and the vertex positions");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N;
  mtg::EvalParams p = {};
  p.K = K; p.N = N; p.derivative = 0; p.max_samples = max_samples; p.v_max = v_max; p.a_max = a_max;
  if (desc->memory == MTG_MEM_DEVICE) {
    p.coeffs = coeffs; p.seg_times = seg_times; p.t_start = t_start; p.t_end = t_end; p.dt = dt;
    p.samples = samples; p.flags = flags; p.max_v = max_v; p.max_a = max_a; p.feasible = feasible;
    p.positions = radii ? positions : nullptr; p.radii = radii;
    p.n_samples = n_samples; p.status = status;
    p.B = B; p.b0 = 0; p.nb = B;
    p.vec_ok = ((uintptr_t)coeffs % 16 == 0) ? 1 : 0;
    return launch_feasibility(ctx, D, aos, p, stream);
  }
  const size_t S = (size_t)max_samples;
  std::vector<HostTensor> ts = {
      {coeffs, (size_t)K * D * N, 8, true, false, nullptr}, {seg_times, (size_t)K, 8, true, false, nullptr},
      {t_start, 1, 8, true, true, nullptr}, {t_end, 1, 8, true, true, nullptr}, {dt, 1, 8, true, true, nullptr},
      {radii ? positions : nullptr, (size_t)(K + 1) * 3, 8, true, false, nullptr},
      {radii, (size_t)K * 2, 8, true, false, nullptr},
      {S ? samples : nullptr, S * D, 8, false, false, nullptr}, {S ? flags : nullptr, S, 1, false, false, nullptr},
      {max_v, 1, 8, false, true, nullptr}, {max_a, 1, 8, false, true, nullptr},
      {feasible, 1, 1, false, true, nullptr}, {n_samples, 1, 4, false, true, nullptr},
      {status, 1, 4, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    p.coeffs = (const double*)ts[0].dev; p.seg_times = (const double*)ts[1].dev;
    p.t_start = (const double*)ts[2].dev; p.t_end = (const double*)ts[3].dev; p.dt = (const double*)ts[4].dev;
    p.positions = (const double*)ts[5].dev; p.radii = (const double*)ts[6].dev;
    p.samples = (double*)ts[7].dev; p.flags = (uint8_t*)ts[8].dev;
    p.max_v = (double*)ts[9].dev; p.max_a = (double*)ts[10].dev; p.feasible = (uint8_t*)ts[11].dev;
    p.n_samples = (int32_t*)ts[12].dev; p.status = (uint32_t*)ts[13].dev;
    p.B = C; p.b0 = 0; p.nb = nb; p.vec_ok = 1;
    return launch_feasibility(ctx, D, aos, p, st);
  });
}

}  // extern "C"

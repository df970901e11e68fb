// libmtg_cuda.so — P9: mtg_cost_time_fd_batch (finite-difference segment-time perturbations).
#include "host_common.h"
#include "cost_fd.cuh"

MTG_REGISTER_TABLES()

using namespace mtg;

namespace {

template <int HN, int D, bool AOS>
int launch_fd_t(mtg_ctx* ctx, const CostFdParams& p, cudaStream_t s) {
  const int block = 128, grid = (p.nb + block - 1) / block;
  if (grid == 0) return MTG_OK;
  cost_time_fd_kernel<HN, D, AOS><<<grid, block, 0, s>>>(p);
  ++ctx->launches;
  MTG_CUDA_TRY(cudaGetLastError());
  return MTG_OK;
}
template <int HN, bool AOS>
int launch_fd_d(mtg_ctx* ctx, int D, const CostFdParams& p, cudaStream_t s) {
  switch (D) {
    case 1: return launch_fd_t<HN, 1, AOS>(ctx, p, s);
    case 2: return launch_fd_t<HN, 2, AOS>(ctx, p, s);
    case 3: return launch_fd_t<HN, 3, AOS>(ctx, p, s);
    case 4: return launch_fd_t<HN, 4, AOS>(ctx, p, s);
  }
  return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "D must be 1..4");
}
template <bool AOS>
int launch_fd_n(mtg_ctx* ctx, int N, int D, const CostFdParams& p, cudaStream_t s) {
  switch (N) {
    case 4: return launch_fd_d<2, AOS>(ctx, D, p, s);
    case 6: return launch_fd_d<3, AOS>(ctx, D, p, s);
    case 8: return launch_fd_d<4, AOS>(ctx, D, p, s);
    case 10: return launch_fd_d<5, AOS>(ctx, D, p, s);
    case 12: return launch_fd_d<6, AOS>(ctx, D, p, s);
  }
  return fail(ctx, MTG_ERR_UNSUPPORTED, "cost_time_fd supports N in {4,6,8,10,12}");
}

template <int HN, int D, bool AOS>
int launch_cf_t(mtg_ctx* ctx, const CostFdParams& p, const SolveCanonicalParams& sp, cudaStream_t s) {
  const int block = 128, grid = (p.nb + block - 1) / block;
  if (grid == 0) return MTG_OK;
  coeffs_from_free_kernel<HN, D, AOS><<<grid, block, 0, s>>>(p, sp);
  ++ctx->launches;
  MTG_CUDA_TRY(cudaGetLastError());
  return MTG_OK;
}
template <int HN, bool AOS>
int launch_cf_d(mtg_ctx* ctx, int D, const CostFdParams& p, const SolveCanonicalParams& sp, cudaStream_t s) {
  switch (D) {
    case 1: return launch_cf_t<HN, 1, AOS>(ctx, p, sp, s);
    case 2: return launch_cf_t<HN, 2, AOS>(ctx, p, sp, s);
    case 3: return launch_cf_t<HN, 3, AOS>(ctx, p, sp, s);
    case 4: return launch_cf_t<HN, 4, AOS>(ctx, p, sp, s);
  }
  return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "D must be 1..4");
}
template <bool AOS>
int launch_cf_n(mtg_ctx* ctx, int N, int D, const CostFdParams& p, const SolveCanonicalParams& sp, cudaStream_t s) {
  switch (N) {
    case 4: return launch_cf_d<2, AOS>(ctx, D, p, sp, s);
    case 6: return launch_cf_d<3, AOS>(ctx, D, p, sp, s);
    case 8: return launch_cf_d<4, AOS>(ctx, D, p, sp, s);
    case 10: return launch_cf_d<5, AOS>(ctx, D, p, sp, s);
    case 12: return launch_cf_d<6, AOS>(ctx, D, p, sp, s);
  }
  return fail(ctx, MTG_ERR_UNSUPPORTED, "supported N: {4,6,8,10,12}");
}

}  // namespace

extern "C" int mtg_set_free_constraints_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* positions,
                                              const double* end_derivatives, const double* seg_times,
                                              const double* free_constraints, double* coeffs, double* cost,
                                              uint32_t* status, void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (!positions || !seg_times || !coeffs)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "positions, seg_times and coeffs are required");
  if (desc->K > 1 && !free_constraints)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "free_constraints (d_p) is required when K > 1");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  TableGuard tables(ctx, desc->N, desc->derivative_to_optimize, (cudaStream_t)stream_);
  if (tables.rc()) return tables.rc();
  cudaStream_t stream = (cudaStream_t)stream_;
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N, NF = N / 2 - 1;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  CostFdParams p = {};
  SolveCanonicalParams sp = {};
  p.K = sp.K = K;
  p.derivative = sp.derivative = desc->derivative_to_optimize;
  auto launch = [&](cudaStream_t st) {
    sp.B = p.B; sp.b0 = p.b0; sp.nb = p.nb;
    return aos ? launch_cf_n<true>(ctx, N, D, p, sp, st) : launch_cf_n<false>(ctx, N, D, p, sp, st);
  };
  if (desc->memory == MTG_MEM_DEVICE) {
    p.positions = positions; p.end_derivatives = end_derivatives; p.seg_times = seg_times;
    p.free_constraints = free_constraints; p.status = status;
    sp.coeffs = coeffs; sp.cost = cost; sp.vec_ok = ((uintptr_t)coeffs % 16 == 0) ? 1 : 0;
    p.B = B; p.b0 = 0; p.nb = B;
    return launch(stream);
  }
  const size_t rec_free = (size_t)std::max(K - 1, 0) * NF * D;
  std::vector<HostTensor> ts = {
      {positions, (size_t)(K + 1) * D, 8, true, false, nullptr},
      {end_derivatives, (size_t)2 * NF * D, 8, true, false, nullptr},
      {seg_times, (size_t)K, 8, true, false, nullptr},
      {rec_free ? free_constraints : nullptr, rec_free, 8, true, false, nullptr},
      {coeffs, (size_t)K * D * N, 8, false, false, nullptr},
      {cost, 1, 8, false, true, nullptr},
      {status, 1, 4, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    p.positions = (const double*)ts[0].dev; p.end_derivatives = (const double*)ts[1].dev;
    p.seg_times = (const double*)ts[2].dev; p.free_constraints = (const double*)ts[3].dev;
    sp.coeffs = (double*)ts[4].dev; sp.cost = (double*)ts[5].dev; p.status = (uint32_t*)ts[6].dev;
    sp.vec_ok = 1;
    p.B = C; p.b0 = 0; p.nb = nb;
    return launch(st);
  });
}

namespace {
template <int HN, int D, bool AOS>
int launch_cd_t(mtg_ctx* ctx, const double* derivs, const SolveCanonicalParams& sp, cudaStream_t s) {
  const int block = 128, grid = (sp.nb + block - 1) / block;
  if (grid == 0) return MTG_OK;
  coeffs_from_derivatives_kernel<HN, D, AOS><<<grid, block, 0, s>>>(derivs, sp);
  ++ctx->launches;
  MTG_CUDA_TRY(cudaGetLastError());
  return MTG_OK;
}
template <int HN, bool AOS>
int launch_cd_d(mtg_ctx* ctx, int D, const double* derivs, const SolveCanonicalParams& sp, cudaStream_t s) {
  switch (D) {
    case 1: return launch_cd_t<HN, 1, AOS>(ctx, derivs, sp, s);
    case 2: return launch_cd_t<HN, 2, AOS>(ctx, derivs, sp, s);
    case 3: return launch_cd_t<HN, 3, AOS>(ctx, derivs, sp, s);
    case 4: return launch_cd_t<HN, 4, AOS>(ctx, derivs, sp, s);
  }
  return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "D must be 1..4");
}
template <bool AOS>
int launch_cd_n(mtg_ctx* ctx, int N, int D, const double* derivs, const SolveCanonicalParams& sp, cudaStream_t s) {
  switch (N) {
    case 4: return launch_cd_d<2, AOS>(ctx, D, derivs, sp, s);
    case 6: return launch_cd_d<3, AOS>(ctx, D, derivs, sp, s);
    case 8: return launch_cd_d<4, AOS>(ctx, D, derivs, sp, s);
    case 10: return launch_cd_d<5, AOS>(ctx, D, derivs, sp, s);
    case 12: return launch_cd_d<6, AOS>(ctx, D, derivs, sp, s);
  }
  return fail(ctx, MTG_ERR_UNSUPPORTED, "supported N: {4,6,8,10,12}");
}
}  // namespace

extern "C" int mtg_coeffs_from_derivatives_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* derivatives,
                                                 const double* seg_times, double* coeffs, double* cost,
                                                 uint32_t* status, void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (!derivatives || !seg_times || !coeffs)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "derivatives, seg_times and coeffs are required");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  TableGuard tables(ctx, desc->N, desc->derivative_to_optimize, (cudaStream_t)stream_);
  if (tables.rc()) return tables.rc();
  cudaStream_t stream = (cudaStream_t)stream_;
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  SolveCanonicalParams sp = {};
  sp.K = K;
  sp.derivative = desc->derivative_to_optimize;
  auto launch = [&](const double* dv, cudaStream_t st) {
    return aos ? launch_cd_n<true>(ctx, N, D, dv, sp, st) : launch_cd_n<false>(ctx, N, D, dv, sp, st);
  };
  if (desc->memory == MTG_MEM_DEVICE) {
    sp.seg_times = seg_times; sp.coeffs = coeffs; sp.cost = cost; sp.status = status;
    sp.B = B; sp.b0 = 0; sp.nb = B; sp.vec_ok = ((uintptr_t)coeffs % 16 == 0) ? 1 : 0;
    return launch(derivatives, stream);
  }
  std::vector<HostTensor> ts = {
      {derivatives, (size_t)(K + 1) * (N / 2) * D, 8, true, false, nullptr},
      {seg_times, (size_t)K, 8, true, false, nullptr},
      {coeffs, (size_t)K * D * N, 8, false, false, nullptr},
      {cost, 1, 8, false, true, nullptr},
      {status, 1, 4, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    sp.seg_times = (const double*)ts[1].dev; sp.coeffs = (double*)ts[2].dev; sp.cost = (double*)ts[3].dev;
    sp.status = (uint32_t*)ts[4].dev;
    sp.B = C; sp.b0 = 0; sp.nb = nb; sp.vec_ok = 1;
    return launch((const double*)ts[0].dev, st);
  });
}

extern "C" int mtg_cost_time_fd_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* positions,
                                      const double* end_derivatives, const double* seg_times,
                                      const double* free_constraints, double increment_time, int central,
                                      double* J_nominal, double* J_plus, double* J_minus, double* grad,
                                      uint32_t* status, void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (!positions || !seg_times) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "positions and seg_times are required");
  if (desc->K > 1 && !free_constraints)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "free_constraints (d_p) is required when K > 1");
  if (!(increment_time > 0.0)) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "increment_time must be > 0");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  TableGuard tables(ctx, desc->N, desc->derivative_to_optimize, (cudaStream_t)stream_);
  if (tables.rc()) return tables.rc();
  cudaStream_t stream = (cudaStream_t)stream_;
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N, NF = N / 2 - 1;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  CostFdParams p = {};
  p.K = K;
  p.derivative = desc->derivative_to_optimize;
  p.increment_time = increment_time;
  p.central = central ? 1 : 0;
  auto launch = [&](cudaStream_t st) {
    return aos ? launch_fd_n<true>(ctx, N, D, p, st) : launch_fd_n<false>(ctx, N, D, p, st);
  };
  if (desc->memory == MTG_MEM_DEVICE) {
    p.positions = positions; p.end_derivatives = end_derivatives; p.seg_times = seg_times;
    p.free_constraints = free_constraints; p.J_nominal = J_nominal; p.J_plus = J_plus;
    p.J_minus = central ? J_minus : nullptr; p.grad = grad; p.status = status;
    p.B = B; p.b0 = 0; p.nb = B;
    return launch(stream);
  }
  const size_t rec_free = (size_t)std::max(K - 1, 0) * NF * D;
  std::vector<HostTensor> ts = {
      {positions, (size_t)(K + 1) * D, 8, true, false, nullptr},
      {end_derivatives, (size_t)2 * NF * D, 8, true, false, nullptr},
      {seg_times, (size_t)K, 8, true, false, nullptr},
      {rec_free ? free_constraints : nullptr, rec_free, 8, true, false, nullptr},
      {J_nominal, 1, 8, false, true, nullptr},
      {J_plus, (size_t)K, 8, false, false, nullptr},
      {central ? J_minus : nullptr, (size_t)K, 8, false, false, nullptr},
      {grad, (size_t)K, 8, false, false, nullptr},
      {status, 1, 4, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    p.positions = (const double*)ts[0].dev; p.end_derivatives = (const double*)ts[1].dev;
    p.seg_times = (const double*)ts[2].dev; p.free_constraints = (const double*)ts[3].dev;
    p.J_nominal = (double*)ts[4].dev; p.J_plus = (double*)ts[5].dev; p.J_minus = (double*)ts[6].dev;
    p.grad = (double*)ts[7].dev; p.status = (uint32_t*)ts[8].dev;
    p.B = C; p.b0 = 0; p.nb = nb;
    return launch(st);
  });
}

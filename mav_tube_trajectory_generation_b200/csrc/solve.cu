// libmtg_cuda.so — P1..P8: mtg_solve_batch (canonical constraint pattern).
#include "host_common.h"
#include "solve_canonical.cuh"

MTG_REGISTER_TABLES()

using namespace mtg;

namespace {

// -------------------------------------------------------------- solve launch
template <int HN, int D, bool AOS, int DT>
int launch_solve_canonical_dt(mtg_ctx* ctx, const mtg::SolveCanonicalParams& p, cudaStream_t stream) {
  constexpr int NF = HN - 1;
  constexpr int SLOTS = NF * NF + NF * D;
  // two lanes per trajectory; each parks (G_j, z_j) of all but the last vertex it eliminates
  const int K = p.K, m = K / 2;
  const int n_own_max = std::max(K - 1 - m, m - 1);
  const size_t per_thread = (size_t)std::max(n_own_max - 1, 0) * SLOTS * sizeof(double);
  const size_t optin = ctx->smem_optin;
  int block = 128;
  if (const char* env = std::getenv("MTG_SOLVE_BLOCK")) {
    block = std::max(2, std::min(128, std::atoi(env))) & ~1;
  } else if (per_thread > 0) {
    const size_t half_sm = (optin + 1024) / 2 - 1024;  // two CTAs per SM, 1 KB reserved each
    if (per_thread * 128 <= half_sm)
      block = 128;
    else if (per_thread * 32 <= optin)
      block = (int)std::min<size_t>(128, (optin / per_thread) / 32 * 32);
    else
      block = (int)(optin / per_thread) & ~1;
  }
  if (block < 2 || per_thread * block > optin)
    return fail(ctx, MTG_ERR_UNSUPPORTED,
                "solve_canonical: K too large for the shared-memory sweep state; use mtg_solve_generic_batch");
  const size_t smem = per_thread * block;
  auto kern = mtg::solve_canonical_kernel<HN, D, AOS, DT>;
  if (smem > 48 * 1024)  // per device and per instantiation; a cheap host-side call
    MTG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)optin));
  const long long threads = 2LL * p.nb;
  const int grid = (int)((threads + block - 1) / block);
  if (grid == 0) return MTG_OK;
  kern<<<grid, block, smem, stream>>>(p);
  ++ctx->launches;
  MTG_CUDA_TRY(cudaGetLastError());
  return MTG_OK;
}

// the default cost derivative N/2 - 1 (kHighestDerivativeToOptimize, LIN_H:51) gets its own instantiation
template <int HN, int D, bool AOS>
int launch_solve_canonical_t(mtg_ctx* ctx, const mtg::SolveCanonicalParams& p, cudaStream_t stream) {
  // (only for N = 10, the reference's default PolynomialOptimization<10>: keeps the build short)
  if (HN == 5 && p.derivative == HN - 1) return launch_solve_canonical_dt<HN, D, AOS, (HN == 5 ? HN - 1 : -1)>(ctx, p, stream);
  return launch_solve_canonical_dt<HN, D, AOS, -1>(ctx, p, stream);
}

template <int HN, bool AOS>
int launch_solve_canonical_d(mtg_ctx* ctx, int D, const mtg::SolveCanonicalParams& p, cudaStream_t s) {
  switch (D) {
    case 1: return launch_solve_canonical_t<HN, 1, AOS>(ctx, p, s);
    case 2: return launch_solve_canonical_t<HN, 2, AOS>(ctx, p, s);
    case 3: return launch_solve_canonical_t<HN, 3, AOS>(ctx, p, s);
    case 4: return launch_solve_canonical_t<HN, 4, AOS>(ctx, p, s);
  }
  return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "D must be 1..4");
}

template <bool AOS>
int launch_solve_canonical_n(mtg_ctx* ctx, int N, int D, const mtg::SolveCanonicalParams& p, cudaStream_t s) {
  switch (N) {
    case 4: return launch_solve_canonical_d<2, AOS>(ctx, D, p, s);
    case 6: return launch_solve_canonical_d<3, AOS>(ctx, D, p, s);
    case 8: return launch_solve_canonical_d<4, AOS>(ctx, D, p, s);
    case 10: return launch_solve_canonical_d<5, AOS>(ctx, D, p, s);
    case 12: return launch_solve_canonical_d<6, AOS>(ctx, D, p, s);
  }
  return fail(ctx, MTG_ERR_UNSUPPORTED, "solve_canonical supports N in {4,6,8,10,12}");
}

int launch_solve_canonical(mtg_ctx* ctx, int N, int D, bool aos, const mtg::SolveCanonicalParams& p,
                           cudaStream_t s) {
  return aos ? launch_solve_canonical_n<true>(ctx, N, D, p, s) : launch_solve_canonical_n<false>(ctx, N, D, p, s);
}

}  // namespace

extern "C" {

int mtg_solve_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* positions,
                    const double* end_derivatives, const double* seg_times, double* coeffs,
                    double* cost, double* free_constraints, uint32_t* status, void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (!positions || !seg_times || !coeffs)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "positions, seg_times and coeffs are required");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  rc = ensure_tables(ctx, desc->N, desc->derivative_to_optimize);
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N, NF = N / 2 - 1;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;

  mtg::SolveCanonicalParams p;
  p.K = K;
  p.derivative = desc->derivative_to_optimize;
  if (desc->memory == MTG_MEM_DEVICE) {
    p.positions = positions;
    p.end_derivatives = end_derivatives;
    p.seg_times = seg_times;
    p.coeffs = coeffs;
    p.cost = cost;
    p.free_constraints = free_constraints;
    p.status = status;
    p.B = B;
    p.b0 = 0;
    p.nb = B;
    p.vec_ok = ((uintptr_t)coeffs % 16 == 0) ? 1 : 0;
    return launch_solve_canonical(ctx, N, D, aos, p, stream);
  }

  // ---- host-memory mode
  const size_t rec_pos = (size_t)(K + 1) * D, rec_end = (size_t)2 * NF * D, rec_t = K;
  const size_t rec_c = (size_t)K * D * N, rec_free = (size_t)std::max(K - 1, 0) * NF * D;
  std::vector<HostTensor> ts = {
      {positions, rec_pos, 8, true, false, nullptr},  {end_derivatives, rec_end, 8, true, false, nullptr},
      {seg_times, rec_t, 8, true, false, nullptr},    {coeffs, rec_c, 8, false, false, nullptr},
      {cost, 1, 8, false, true, nullptr},             {rec_free ? free_constraints : nullptr, rec_free, 8, false, false, nullptr},
      {status, 1, 4, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    p.positions = (const double*)ts[0].dev;
    p.end_derivatives = (const double*)ts[1].dev;
    p.seg_times = (const double*)ts[2].dev;
    p.coeffs = (double*)ts[3].dev;
    p.cost = (double*)ts[4].dev;
    p.free_constraints = (double*)ts[5].dev;
    p.status = (uint32_t*)ts[6].dev;
    p.B = C;
    p.b0 = 0;
    p.nb = nb;
    p.vec_ok = 1;
    return launch_solve_canonical(ctx, N, D, aos, p, st);
  });
}

}  // extern "C"

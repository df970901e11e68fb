// libmtg_cuda.so — mtg_solve_generic_batch: arbitrary (batch-shared) constraint pattern.
#include "host_common.h"
#include "solve_generic.cuh"

MTG_REGISTER_TABLES()

using namespace mtg;

namespace {
constexpr int kGenericChunk = 32768;  // trajectories per launch (bounds the parked-factor scratch)

template <int HN, int D, bool AOS>
int launch_generic_t(mtg_ctx* ctx, const SolveCanonicalParams& p_in, const SolveGenericParams& g_in, const uint8_t* mask_host,
                     cudaStream_t s) {
  constexpr int SL = HN * HN + HN * D;
  SolveCanonicalParams p = p_in;
  SolveGenericParams g = g_in;
  const int chunk = std::min(p_in.nb, kGenericChunk);
  const size_t mask_bytes = align256((size_t)(p.K + 1) * HN);
  DeviceBuffer* scratch = ctx->scratch_for(s);
  if (scratch->ensure(mask_bytes + (size_t)(p.K + 1) * SL * chunk * sizeof(double)))
    return fail(ctx, MTG_ERR_CUDA, "cudaMalloc of the generic-solve scratch failed");
  MTG_CUDA_TRY(cudaMemcpyAsync(scratch->ptr, mask_host, (size_t)(p.K + 1) * HN, cudaMemcpyHostToDevice, s));
  g.mask = (const uint8_t*)scratch->ptr;
  g.scratch = (double*)((char*)scratch->ptr + mask_bytes);
  for (int off = 0; off < p_in.nb; off += chunk) {
    p.b0 = p_in.b0 + off;
    p.nb = std::min(chunk, p_in.nb - off);
    solve_generic_kernel<HN, D, AOS><<<(p.nb + 127) / 128, 128, 0, s>>>(p, g);
    ++ctx->launches;
    MTG_CUDA_TRY(cudaGetLastError());
  }
  return MTG_OK;
}
template <int HN, bool AOS>
int launch_generic_d(mtg_ctx* ctx, int D, const SolveCanonicalParams& p, const SolveGenericParams& g, const uint8_t* m,
                     cudaStream_t s) {
  switch (D) {
    case 1: return launch_generic_t<HN, 1, AOS>(ctx, p, g, m, s);
    case 2: return launch_generic_t<HN, 2, AOS>(ctx, p, g, m, s);
    case 3: return launch_generic_t<HN, 3, AOS>(ctx, p, g, m, s);
    case 4: return launch_generic_t<HN, 4, AOS>(ctx, p, g, m, s);
  }
  return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "D must be 1..4");
}
template <bool AOS>
int launch_generic_n(mtg_ctx* ctx, int N, int D, const SolveCanonicalParams& p, const SolveGenericParams& g,
                     const uint8_t* m, cudaStream_t s) {
  switch (N) {
    case 4: return launch_generic_d<2, AOS>(ctx, D, p, g, m, s);
    case 6: return launch_generic_d<3, AOS>(ctx, D, p, g, m, s);
    case 8: return launch_generic_d<4, AOS>(ctx, D, p, g, m, s);
    case 10: return launch_generic_d<5, AOS>(ctx, D, p, g, m, s);
    case 12: return launch_generic_d<6, AOS>(ctx, D, p, g, m, s);
  }
  return fail(ctx, MTG_ERR_UNSUPPORTED, "supported N: {4,6,8,10,12}");
}
}  // namespace

extern "C" int mtg_solve_generic_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const uint8_t* mask,
                                       const double* values, const double* seg_times, double* coeffs, double* cost,
                                       double* free_constraints, uint32_t* status, void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (!mask || !values || !seg_times || !coeffs)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "mask, values, seg_times and coeffs are required");
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N, h = N / 2;
  int n_free = 0;
  for (int i = 0; i < (K + 1) * h; ++i) n_free += mask[i] ? 0 : 1;
  if (B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  TableGuard tables(ctx, N, desc->derivative_to_optimize, (cudaStream_t)stream_);
  if (tables.rc()) return tables.rc();
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  SolveCanonicalParams p = {};
  SolveGenericParams g = {};
  p.K = K;
  p.derivative = desc->derivative_to_optimize;
  g.n_free = n_free;
  auto launch = [&](cudaStream_t st) {
    return aos ? launch_generic_n<true>(ctx, N, D, p, g, mask, st) : launch_generic_n<false>(ctx, N, D, p, g, mask, st);
  };
  if (desc->memory == MTG_MEM_DEVICE) {
    p.seg_times = seg_times; p.coeffs = coeffs; p.cost = cost; p.status = status;
    g.values = values; g.free_out = free_constraints;
    p.B = B; p.b0 = 0; p.nb = B;
    p.vec_ok = ((uintptr_t)coeffs % 16 == 0) ? 1 : 0;
    return launch(stream);
  }
  std::vector<HostTensor> ts = {
      {values, (size_t)(K + 1) * h * D, 8, true, false, nullptr},
      {seg_times, (size_t)K, 8, true, false, nullptr},
      {coeffs, (size_t)K * D * N, 8, false, false, nullptr},
      {cost, 1, 8, false, true, nullptr},
      {n_free ? free_constraints : nullptr, (size_t)D * n_free, 8, false, false, nullptr},
      {status, 1, 4, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    g.values = (const double*)ts[0].dev; p.seg_times = (const double*)ts[1].dev;
    p.coeffs = (double*)ts[2].dev; p.cost = (double*)ts[3].dev; g.free_out = (double*)ts[4].dev;
    p.status = (uint32_t*)ts[5].dev;
    p.B = C; p.b0 = 0; p.nb = nb; p.vec_ok = 1;
    return launch(st);
  });
}

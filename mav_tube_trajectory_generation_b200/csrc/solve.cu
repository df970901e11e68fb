// libmtg_cuda.so — P1..P8: mtg_solve_batch (canonical constraint pattern).
#include "host_common.h"
#include "solve_canonical.cuh"


using namespace mtg;

namespace mtg {
// solve_soa.cu / solve_aos.cu: one translation unit per layout (they compile in parallel)
int launch_solve_canonical_soa(mtg_ctx* ctx, int N, int D, const SolveCanonicalParams& p, cudaStream_t s);
int launch_solve_canonical_aos(mtg_ctx* ctx, int N, int D, const SolveCanonicalParams& p, cudaStream_t s);
}  // namespace mtg

namespace {
int launch_solve_canonical(mtg_ctx* ctx, int N, int D, bool aos, const mtg::SolveCanonicalParams& p,
                           cudaStream_t s) {
  return aos ? mtg::launch_solve_canonical_aos(ctx, N, D, p, s) : mtg::launch_solve_canonical_soa(ctx, N, D, p, s);
}
}  // namespace

extern "C" {

int mtg_solve_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* positions,
                    const double* end_derivatives, const double* seg_times, double* coeffs,
                    double* cost, double* free_constraints, uint32_t* status, void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (!positions || !seg_times) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "positions and seg_times are required");
  if (!coeffs && !cost && !free_constraints)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "at least one of coeffs, cost and free_constraints is required");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  TableGuard tables(ctx, desc->N, desc->derivative_to_optimize, (cudaStream_t)stream_);
  if (tables.rc()) return tables.rc();
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
    p.vec_ok = (coeffs && (uintptr_t)coeffs % 16 == 0) ? 1 : 0;
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

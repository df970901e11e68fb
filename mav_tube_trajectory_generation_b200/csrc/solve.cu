// libmtg_cuda.so — P1..P8: mtg_solve_batch (canonical constraint pattern).
#include "host_common.h"
#include "solve_canonical.cuh"
#include "solve_launch.cuh"  // solve_block_size()


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
    p.overlap = ctx->solve_overlap;
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

int mtg_set_solve_overlap(mtg_ctx* ctx, int enabled) {
  if (!ctx) return MTG_ERR_INVALID_ARGUMENT;
  ctx->solve_overlap = enabled ? 1 : 0;
  return MTG_OK;
}

// Candidate sweep in ONE launch: the solve with the argmin of its costs folded into the kernel's epilogue.
int mtg_solve_argmin_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* positions,
                           const double* end_derivatives, const double* seg_times, double* coeffs, double* cost,
                           double* free_constraints, uint32_t* status, int64_t global_offset, int accumulate,
                           void* best, void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (desc->memory != MTG_MEM_DEVICE)
    return fail(ctx, MTG_ERR_UNSUPPORTED, "mtg_solve_argmin_batch takes device pointers (the running best lives on the device)");
  if (!positions || !seg_times || !best)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "positions, seg_times and best are required");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N;
  if (B == 0) return mtg_argmin_batch(ctx, nullptr, nullptr, 0, global_offset, accumulate, best, stream_);
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  const int block = mtg::solve_launch::solve_block_size(ctx, K, N / 2, D, nullptr);
  if (K < 2 || block < 32 || block % 32 != 0) {
    // shapes the fused epilogue does not cover (a single segment; very long chains that run in partial warps):
    // the two launches it replaces, with scratch for whatever the caller did not ask for
    double* cost_buf = cost;
    uint32_t* status_buf = status;
    if (!cost_buf || !status_buf) {
      DeviceBuffer* scratch = ctx->scratch_for(stream);
      const size_t cost_bytes = align256((size_t)B * sizeof(double));
      if (scratch->ensure(cost_bytes + (size_t)B * sizeof(uint32_t)))
        return fail(ctx, MTG_ERR_CUDA, "cudaMalloc of the sweep scratch failed");
      if (!cost_buf) cost_buf = (double*)scratch->ptr;
      if (!status_buf) status_buf = (uint32_t*)((char*)scratch->ptr + cost_bytes);
    }
    rc = mtg_solve_batch(ctx, desc, positions, end_derivatives, seg_times, coeffs, cost_buf, free_constraints,
                         status_buf, stream_);
    if (rc) return rc;
    return mtg_argmin_batch(ctx, cost_buf, status_buf, B, global_offset, accumulate, best, stream_);
  }
  TableGuard tables(ctx, N, desc->derivative_to_optimize, stream);
  if (tables.rc()) return tables.rc();
  DeviceBuffer* state = ctx->argmin_state_for(stream);  // [ticket, lock | partials]: shared with mtg_argmin_batch
  if (!state->ptr) {
    const size_t need = 256 + (size_t)4 * std::max(ctx->sm_count, 1) * sizeof(mtg::Best);
    if (state->ensure(need)) return fail(ctx, MTG_ERR_CUDA, "cudaMalloc of the argmin state failed");
    MTG_CUDA_TRY(cudaMemset(state->ptr, 0, 256));  // ticket counter and lock, once: the kernels leave them at zero
  }
  if (!accumulate) {  // a fresh sweep: the pair starts as "nothing yet"
    rc = mtg_argmin_batch(ctx, nullptr, nullptr, 0, 0, 0, best, stream_);
    if (rc) return rc;
  }
  mtg::SolveCanonicalParams p;
  p.K = K;
  p.derivative = desc->derivative_to_optimize;
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
  p.overlap = ctx->solve_overlap;
  p.best_lock = (unsigned*)state->ptr + 1;
  p.best_out = best;
  p.best_offset = global_offset;
  return launch_solve_canonical(ctx, N, D, desc->layout == MTG_LAYOUT_AOS, p, stream);
}

}  // extern "C"

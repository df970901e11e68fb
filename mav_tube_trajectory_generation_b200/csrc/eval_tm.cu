// libmtg_cuda.so — launchers of the time-major sweep kernels (eval_tm.cuh).
#include <cmath>

#include "host_common.h"
#include "eval_tm.cuh"

MTG_REGISTER_BASE()

using namespace mtg;

namespace {

// largest double y with sqrt(y) <= lim (sqrt correctly rounded and monotone), so that
// (|v|^2 <= y) == (sqrt(|v|^2) <= lim) for every |v|^2 >= 0
double sq_limit(double lim) {
  if (lim != lim) return lim;  // NaN: every comparison is false
  if (lim < 0.0) return -1.0;
  if (std::isinf(lim)) return lim;
  double y = lim * lim;
  while (std::sqrt(y) > lim) y = std::nextafter(y, -INFINITY);
  while (std::sqrt(std::nextafter(y, INFINITY)) <= lim) y = std::nextafter(y, INFINITY);
  return y;
}

constexpr int kTmBlock = 32;            // one warp per CTA = tm_tpw() = 16 trajectories (finest shared-memory packing)
constexpr int kTubeChunk = 65536;       // trajectories per launch pair when a tube scratch is needed

template <int NT, int D, int MODE>
int launch_tm_t(mtg_ctx* ctx, const EvalParams& p, const double* geom, cudaStream_t s) {
  const bool want_acc = MODE == TM_DERIVATIVE && p.sampling_times != nullptr;
  const size_t smem = (size_t)tm_layout(D, NT, want_acc, MODE == TM_FEAS_TUBE, tm_tpw(MODE), MODE).per_warp * (kTmBlock / 32);
  auto kern = eval_tm_kernel<NT, D, MODE>;
  const int per_block = (kTmBlock / 32) * tm_tpw(MODE);  // trajectories per CTA
  if (smem > 48 * 1024) MTG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (p.nb + per_block - 1) / per_block;
  if (grid == 0) return MTG_OK;
  kern<<<grid, kTmBlock, smem, s>>>(p, geom);
  ++ctx->launches;
  MTG_CUDA_TRY(cudaGetLastError());
  return MTG_OK;
}

template <int NT, int MODE>
int launch_tm_d(mtg_ctx* ctx, int D, const EvalParams& p, const double* geom, cudaStream_t s) {
  switch (D) {
    case 1: return launch_tm_t<NT, 1, MODE>(ctx, p, geom, s);
    case 2: return launch_tm_t<NT, 2, MODE>(ctx, p, geom, s);
    case 3: return launch_tm_t<NT, 3, MODE>(ctx, p, geom, s);
    case 4: return launch_tm_t<NT, 4, MODE>(ctx, p, geom, s);
  }
  return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "D must be 1..4");
}

template <int MODE>
int launch_tm_n(mtg_ctx* ctx, int D, const EvalParams& p, const double* geom, cudaStream_t s) {
  return p.N == 10 ? launch_tm_d<10, MODE>(ctx, D, p, geom, s) : launch_tm_d<12, MODE>(ctx, D, p, geom, s);
}

}  // namespace

namespace mtg {

bool eval_tm_supported(const EvalParams& p) {
  return (p.N == 10 || p.N == 12) && p.K <= (1 << 30) && ((uintptr_t)p.coeffs % 16 == 0) &&
         ((uintptr_t)p.seg_times % 8 == 0);
}

int launch_eval_tm(mtg_ctx* ctx, int D, bool feasibility, const EvalParams& p_in, cudaStream_t s) {
  EvalParams p = p_in;
  if (!feasibility)
    // the lean position kernel has no sampling_times / segment_idx outputs; TM_DERIVATIVE handles every
    // derivative order (B(0, j) = 1: bit-identical values for order 0)
    return (p.derivative == 0 && !p.sampling_times && !p.segment_idx) ? launch_tm_n<TM_POSITION>(ctx, D, p, nullptr, s)
                             : launch_tm_n<TM_DERIVATIVE>(ctx, D, p, nullptr, s);
  p.v2_lim = sq_limit(p.v_max);
  p.a2_lim = sq_limit(p.a_max);
  const bool tube = D == 3 && p.radii != nullptr && p.positions != nullptr;
  if (!tube) return launch_tm_n<TM_FEAS>(ctx, D, p, nullptr, s);
  // tube constants per (trajectory, segment) go through a scratch buffer, kTubeChunk trajectories at a time
  const int chunk = std::min(p_in.nb, kTubeChunk);
  DeviceBuffer* scratch = ctx->scratch_for(s);
  if (scratch->ensure((size_t)chunk * p.K * kTubeGeomLd * sizeof(double)))
    return fail(ctx, MTG_ERR_CUDA, "cudaMalloc of the tube scratch failed");
  double* geom = (double*)scratch->ptr;
  for (int off = 0; off < p_in.nb; off += chunk) {
    p.b0 = p_in.b0 + off;
    p.nb = std::min(chunk, p_in.nb - off);
    const long long threads = (long long)p.nb * p.K;
    tube_setup_kernel<true><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(p, geom);
    ++ctx->launches;
    MTG_CUDA_TRY(cudaGetLastError());
    const int rc = p.N == 10 ? launch_tm_t<10, 3, TM_FEAS_TUBE>(ctx, p, geom, s)
                             : launch_tm_t<12, 3, TM_FEAS_TUBE>(ctx, p, geom, s);
    if (rc) return rc;
  }
  return MTG_OK;
}

}  // namespace mtg

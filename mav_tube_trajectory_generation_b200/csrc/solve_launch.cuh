// Launchers of solve_canonical_kernel, shared by the per-layout translation units.
#ifndef MTG_SOLVE_LAUNCH_CUH_
#define MTG_SOLVE_LAUNCH_CUH_

#include "host_common.h"
#include "solve_canonical.cuh"

namespace mtg {
namespace solve_launch {

// -------------------------------------------------------------- solve launch
// threads per CTA (2 per trajectory) and the parked sweep state per thread; block < 2: K too large
inline int solve_block_size(const mtg_ctx* ctx, int K, int HN, int D, size_t* per_thread_out) {
  const int NF = HN - 1;
  const int SLOTS = NF * NF + NF * D;
  // two lanes per trajectory; each parks (G_j, z_j) of all but the last vertex it eliminates
  const int m = K / 2;
  const int n_own_max = std::max(K - 1 - m, m - 1);
  const size_t per_thread = (size_t)std::max(n_own_max - 1, 0) * SLOTS * sizeof(double);
  const size_t optin = ctx->smem_optin - 1024;  // the kernel also holds ~0.5 KB of static shared memory
  int block = MTG_SOLVE_THREADS;
  if (const char* env = std::getenv("MTG_SOLVE_BLOCK")) {
    block = std::max(2, std::min(MTG_SOLVE_THREADS, std::atoi(env))) & ~1;
  } else if (per_thread > 0) {
    const size_t half_sm = (optin + 1024) / 2 - 1024;  // two CTAs per SM, 1 KB reserved each
    if (per_thread * MTG_SOLVE_THREADS <= half_sm)
      block = MTG_SOLVE_THREADS;
    else if (per_thread * 32 <= optin)
      block = (int)std::min<size_t>(MTG_SOLVE_THREADS, (optin / per_thread) / 32 * 32);
    else
      block = (int)(optin / per_thread) & ~1;
  }
  if (per_thread_out) *per_thread_out = per_thread;
  return block;
}

template <int HN, int D, bool AOS, int DT>
int launch_solve_canonical_dt(mtg_ctx* ctx, const mtg::SolveCanonicalParams& p, cudaStream_t stream) {
  size_t per_thread = 0;
  const int block = solve_block_size(ctx, p.K, HN, D, &per_thread);
  const size_t optin = ctx->smem_optin - 1024;
  if (block < 2 || per_thread * block > optin)
    return fail(ctx, MTG_ERR_UNSUPPORTED,
                "solve_canonical: K too large for the shared-memory sweep state; use mtg_solve_generic_batch");
  const size_t smem = per_thread * block;
  auto kern = mtg::solve_canonical_kernel<HN, D, AOS, DT>;
  if (smem > 48 * 1024)  // per device and per instantiation; a cheap host-side call
    MTG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long threads = 2LL * p.nb;
  const int grid = (int)((threads + block - 1) / block);
  if (grid == 0) return MTG_OK;
  if (p.best_out && (block % 32 != 0 || p.K < 2))
    return fail(ctx, MTG_ERR_UNSUPPORTED, "solve_canonical: the fused argmin needs whole warps and K >= 2");
  if (p.overlap) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MTG_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
  } else {
    kern<<<grid, block, smem, stream>>>(p);
  }
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

}  // namespace solve_launch
}  // namespace mtg
#endif

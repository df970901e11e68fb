// libmtg_cuda.so — measured roofs for the bench line. MEASURED_PEAKS.json (driver-written) carries the HBM
// copy bandwidth and the bf16 tensor throughput but no fp64 number, and the path computes in fp64 on the
// plain DFMA pipe (there is no fp64 tcgen05 path), so the fp64 roof is measured here: a register-only
// DFMA micro-benchmark, timed with CUDA events on the caller's stream.
#include "host_common.h"

using namespace mtg;

namespace {
constexpr int kChains = 8;

__global__ void __launch_bounds__(256) dfma_probe_kernel(double* __restrict__ out, const double* __restrict__ in,
                                                         int iters) {
  // kChains independent chains per thread: with >= 8 warps per SM sub-partition the issue slots of the
  // fp64 pipe are always covered, so the result is the pipe's throughput, not its latency
  double a[kChains];
  const double x = in[0], y = in[1];
#pragma unroll
  for (int q = 0; q < kChains; ++q) a[q] = (double)(threadIdx.x + q);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < kChains; ++q) a[q] = fma(a[q], x, y);
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < kChains; ++q) s += a[q];
  if (s == 123.456) out[0] = s;  // never true for the inputs used; keeps the chains alive
}
}  // namespace

extern "C" int mtg_probe_fp64_fma(mtg_ctx* ctx, int reps, double* tflops, double* ms_per_launch, void* stream_) {
  if (!ctx || !tflops) return MTG_ERR_INVALID_ARGUMENT;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream_;
  if (ctx->scratch.ensure(256)) return fail(ctx, MTG_ERR_CUDA, "cudaMalloc failed");
  double* buf = (double*)ctx->scratch.ptr;
  const double h[2] = {0.999999, 1e-6};
  MTG_CUDA_TRY(cudaMemcpyAsync(buf + 1, h, sizeof(h), cudaMemcpyHostToDevice, s));
  const int iters = 1 << 15;
  const int grid = std::max(ctx->sm_count, 1) * 8, block = 256;
  cudaEvent_t e0, e1;
  MTG_CUDA_TRY(cudaEventCreate(&e0));
  MTG_CUDA_TRY(cudaEventCreate(&e1));
  reps = std::max(reps, 1);
  dfma_probe_kernel<<<grid, block, 0, s>>>(buf, buf + 1, iters);  // warm-up (clocks, instruction cache)
  MTG_CUDA_TRY(cudaEventRecord(e0, s));
  for (int r = 0; r < reps; ++r) dfma_probe_kernel<<<grid, block, 0, s>>>(buf, buf + 1, iters);
  MTG_CUDA_TRY(cudaEventRecord(e1, s));
  ctx->launches += reps + 1;
  MTG_CUDA_TRY(cudaGetLastError());
  MTG_CUDA_TRY(cudaEventSynchronize(e1));
  float ms = 0.f;
  MTG_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double flop = 2.0 * kChains * (double)iters * (double)grid * block * reps;
  *tflops = flop / (ms * 1e-3) / 1e12;
  if (ms_per_launch) *ms_per_launch = ms / reps;
  return MTG_OK;
}

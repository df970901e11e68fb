// libmtg_cuda.so — E6 / R1: mtg_extrema_batch (analytic extrema of the derivative magnitude).
#include "host_common.h"
#include "extrema.cuh"

MTG_REGISTER_BASE()

using namespace mtg;

namespace {
constexpr int kExtremaChunk = 1 << 18;  // trajectories per launch pair (bounds the scratch: 40 B x K each)

int launch_extrema(mtg_ctx* ctx, bool aos, const ExtremaParams& p_in, cudaStream_t s) {
  ExtremaParams p = p_in;
  const int chunk = std::min(p_in.nb, kExtremaChunk);
  DeviceBuffer* scratch = ctx->scratch_for(s);
  const size_t out_bytes = align256((size_t)chunk * p.K * 4 * sizeof(double));
  if (scratch->ensure(out_bytes + (size_t)chunk * p.K * sizeof(uint32_t)))
    return fail(ctx, MTG_ERR_CUDA, "cudaMalloc of the extrema scratch failed");
  p.seg_out = (double*)scratch->ptr;
  p.seg_status = (uint32_t*)((char*)scratch->ptr + out_bytes);
  for (int off = 0; off < p_in.nb; off += chunk) {
    p.b0 = p_in.b0 + off;
    p.nb = std::min(chunk, p_in.nb - off);
    const long long threads = (long long)p.nb * p.K;
    const unsigned grid = (unsigned)((threads + 127) / 128);
    if (aos)
      extrema_segment_kernel<true><<<grid, 128, 0, s>>>(p);
    else
      extrema_segment_kernel<false><<<grid, 128, 0, s>>>(p);
    ++ctx->launches;
    MTG_CUDA_TRY(cudaGetLastError());
    if (aos)
      extrema_reduce_kernel<true><<<(p.nb + 255) / 256, 256, 0, s>>>(p);
    else
      extrema_reduce_kernel<false><<<(p.nb + 255) / 256, 256, 0, s>>>(p);
    ++ctx->launches;
    MTG_CUDA_TRY(cudaGetLastError());
  }
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
  // LIN_I:400-401 CHECK(N - derivative - 1 > 0)
  if (derivative < 0 || desc->N - derivative - 1 <= 0)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "derivative must satisfy 0 <= derivative < N - 1");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N;
  ExtremaParams p = {};
  p.K = K; p.N = N; p.D = D; p.derivative = derivative;
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

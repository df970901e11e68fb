// libmtg_cuda.so — single CUDA translation unit: context, constant tables,
// kernel launchers and the C ABI of include/mtg_cuda.h.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 (build.py).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mtg_cuda.h"
#include "device_tables.cuh"

#include "solve_canonical.cuh"

namespace {

constexpr int kStageSlots = 3;

struct DeviceBuffer {
  void* ptr = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
    if (cudaMalloc(&ptr, need) != cudaSuccess) return -1;
    bytes = need;
    return 0;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
  }
};

}  // namespace

struct mtg_ctx {
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  std::string err;
  uint64_t launches = 0;
  cudaStream_t stage_stream[kStageSlots] = {nullptr, nullptr, nullptr};
  // per staging slot: inputs / outputs of one chunk (HOST-memory mode)
  DeviceBuffer stage_in[kStageSlots];
  DeviceBuffer stage_out[kStageSlots];
};

namespace {

// (N, derivative) currently resident in c_tab, per device
std::mutex g_tab_mutex;
int g_tab_N[64];
int g_tab_d[64];
bool g_tab_init = false;

int fail(mtg_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

int cuda_fail(mtg_ctx* ctx, cudaError_t e, const char* where) {
  return fail(ctx, MTG_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}

#define MTG_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return cuda_fail(ctx, _e, #expr); \
  } while (0)

int ensure_tables(mtg_ctx* ctx, int N, int derivative) {
  std::lock_guard<std::mutex> lock(g_tab_mutex);
  if (!g_tab_init) {
    for (int i = 0; i < 64; ++i) g_tab_N[i] = g_tab_d[i] = -1;
    g_tab_init = true;
  }
  const int dev = ctx->device & 63;
  if (g_tab_N[dev] == N && g_tab_d[dev] == derivative) return MTG_OK;
  mtg::Tables t;
  if (!mtg::compute_tables(N, derivative, &t))
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "invalid (N, derivative_to_optimize)");
  mtg::DevTables h;
  std::memcpy(h.H1, t.H1, sizeof(h.H1));
  std::memcpy(h.Ainv1, t.Ainv1, sizeof(h.Ainv1));
  std::memcpy(h.W, t.W, sizeof(h.W));
  std::memcpy(h.base, t.base, sizeof(h.base));
  for (int j = 0; j < MTG_TAB_LD; ++j) h.inv_factorial[j] = 1.0 / t.base[j * MTG_BASE_LD + j];
  h.N = N;
  h.derivative = derivative;
  // a config switch is rare: drain the device so no in-flight kernel sees a torn table
  MTG_CUDA_TRY(cudaDeviceSynchronize());
  MTG_CUDA_TRY(cudaMemcpyToSymbol(mtg::c_tab, &h, sizeof(h)));
  g_tab_N[dev] = N;
  g_tab_d[dev] = derivative;
  return MTG_OK;
}

int validate_desc(mtg_ctx* ctx, const mtg_problem_desc* d) {
  if (!ctx) return MTG_ERR_INVALID_ARGUMENT;
  if (!d) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "desc is NULL");
  if (d->B < 0 || d->K < 1) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "B >= 0 and K >= 1 required");
  if (d->D < 1 || d->D > 4) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "D must be 1..4");
  if (d->N < 2 || d->N > MTG_MAX_N || (d->N & 1))
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "N must be even and <= 12");
  // LIN_I:50-55 CHECK(derivative_to_optimize >= 0 && <= kHighestDerivativeToOptimize)
  if (d->derivative_to_optimize < 0 || d->derivative_to_optimize > d->N / 2 - 1)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "derivative_to_optimize must be in [0, N/2-1]");
  if (d->memory != MTG_MEM_DEVICE && d->memory != MTG_MEM_HOST)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "desc.memory must be MTG_MEM_DEVICE or MTG_MEM_HOST");
  if (d->layout != MTG_LAYOUT_SOA && d->layout != MTG_LAYOUT_AOS)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "desc.layout must be MTG_LAYOUT_SOA or MTG_LAYOUT_AOS");
  return MTG_OK;
}

// -------------------------------------------------------------- solve launch
template <int HN, int D, bool AOS>
int launch_solve_canonical_t(mtg_ctx* ctx, const mtg::SolveCanonicalParams& p, cudaStream_t stream) {
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
  auto kern = mtg::solve_canonical_kernel<HN, D, AOS>;
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

// Copies of a chunk [b0, b0+nb) of a batched tensor with `rec` elements per record
// between the caller's host tensor (batch B) and a chunk-sized device tensor (batch C).
cudaError_t h2d_chunk(void* dst, size_t C, const void* src, size_t B, size_t b0, size_t nb, size_t rec,
                      size_t elem, bool aos, cudaStream_t s) {
  if (aos)
    return cudaMemcpyAsync(dst, (const char*)src + b0 * rec * elem, nb * rec * elem, cudaMemcpyHostToDevice, s);
  return cudaMemcpy2DAsync(dst, C * elem, (const char*)src + b0 * elem, B * elem, nb * elem, rec,
                           cudaMemcpyHostToDevice, s);
}
cudaError_t d2h_chunk(void* dst, size_t B, size_t b0, const void* src, size_t C, size_t nb, size_t rec,
                      size_t elem, bool aos, cudaStream_t s) {
  if (aos)
    return cudaMemcpyAsync((char*)dst + b0 * rec * elem, src, nb * rec * elem, cudaMemcpyDeviceToHost, s);
  return cudaMemcpy2DAsync((char*)dst + b0 * elem, B * elem, src, C * elem, nb * elem, rec,
                           cudaMemcpyDeviceToHost, s);
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// ================================================================== C ABI
extern "C" {

int mtg_abi_version(void) { return MTG_ABI_VERSION; }

int mtg_create(int device, mtg_ctx** out) {
  if (!out) return MTG_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return MTG_ERR_NO_DEVICE;
  if (device < 0 || device >= n) return MTG_ERR_INVALID_ARGUMENT;
  if (cudaSetDevice(device) != cudaSuccess) return MTG_ERR_CUDA;
  mtg_ctx* ctx = new mtg_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete ctx;
    return MTG_ERR_CUDA;
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  for (int i = 0; i < kStageSlots; ++i)
    if (cudaStreamCreateWithFlags(&ctx->stage_stream[i], cudaStreamNonBlocking) != cudaSuccess) {
      delete ctx;
      return MTG_ERR_CUDA;
    }
  *out = ctx;
  return MTG_OK;
}

void mtg_destroy(mtg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < kStageSlots; ++i) {
    ctx->stage_in[i].release();
    ctx->stage_out[i].release();
    if (ctx->stage_stream[i]) cudaStreamDestroy(ctx->stage_stream[i]);
  }
  delete ctx;
}

const char* mtg_last_error(const mtg_ctx* ctx) { return ctx ? ctx->err.c_str() : "ctx is NULL"; }

uint64_t mtg_launch_count(const mtg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mtg_sync(mtg_ctx* ctx, void* stream) {
  if (!ctx) return MTG_ERR_INVALID_ARGUMENT;
  MTG_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return MTG_OK;
}

int mtg_get_tables(int N, int derivative, double* H1, double* Ainv1) {
  mtg::Tables t;
  if (!mtg::compute_tables(N, derivative, &t)) return MTG_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      if (H1) H1[i * N + j] = t.H1[i * MTG_TAB_LD + j];
      if (Ainv1) Ainv1[i * N + j] = t.Ainv1[i * MTG_TAB_LD + j];
    }
  return MTG_OK;
}

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

  // ---- host-memory mode: pipelined chunks over kStageSlots streams
  const size_t rec_pos = (size_t)(K + 1) * D, rec_end = (size_t)2 * NF * D, rec_t = K;
  const size_t rec_c = (size_t)K * D * N, rec_free = (size_t)std::max(K - 1, 0) * NF * D;
  const size_t in_rows = rec_pos + (end_derivatives ? rec_end : 0) + rec_t;
  const size_t out_rows = rec_c + (cost ? 1 : 0) + (free_constraints ? rec_free : 0);
  size_t C = 8192;
  if (const char* env = std::getenv("MTG_HOST_CHUNK")) C = std::max(1, std::atoi(env));
  C = std::min<size_t>(C, (size_t)B);
  const size_t in_bytes = align256(in_rows * C * 8);
  const size_t out_bytes = align256(out_rows * C * 8) + align256(C * 4);
  const int n_chunks = (int)((B + C - 1) / C);
  const int slots = std::min(kStageSlots, n_chunks);
  for (int s = 0; s < slots; ++s)
    if (ctx->stage_in[s].ensure(in_bytes) || ctx->stage_out[s].ensure(out_bytes))
      return fail(ctx, MTG_ERR_CUDA, "cudaMalloc of staging buffers failed");
  // order after work already queued on the caller's stream
  MTG_CUDA_TRY(cudaStreamSynchronize(stream));
  for (int c = 0; c < n_chunks; ++c) {
    const int s = c % kStageSlots;
    cudaStream_t st = ctx->stage_stream[s];
    const size_t b0 = (size_t)c * C, nb = std::min(C, (size_t)B - b0);
    double* din = (double*)ctx->stage_in[s].ptr;
    double* d_pos = din;
    double* d_end = d_pos + rec_pos * C;
    double* d_t = d_end + (end_derivatives ? rec_end * C : 0);
    double* dout = (double*)ctx->stage_out[s].ptr;
    double* d_c = dout;
    double* d_cost = d_c + rec_c * C;
    double* d_free = d_cost + (cost ? C : 0);
    uint32_t* d_status = (uint32_t*)((char*)dout + align256(out_rows * C * 8));
    MTG_CUDA_TRY(h2d_chunk(d_pos, C, positions, B, b0, nb, rec_pos, 8, aos, st));
    if (end_derivatives) MTG_CUDA_TRY(h2d_chunk(d_end, C, end_derivatives, B, b0, nb, rec_end, 8, aos, st));
    MTG_CUDA_TRY(h2d_chunk(d_t, C, seg_times, B, b0, nb, rec_t, 8, aos, st));
    p.positions = d_pos;
    p.end_derivatives = end_derivatives ? d_end : nullptr;
    p.seg_times = d_t;
    p.coeffs = d_c;
    p.cost = cost ? d_cost : nullptr;
    p.free_constraints = free_constraints ? d_free : nullptr;
    p.status = status ? d_status : nullptr;
    p.B = (int)C;
    p.b0 = 0;
    p.nb = (int)nb;
    p.vec_ok = 1;
    rc = launch_solve_canonical(ctx, N, D, aos, p, st);
    if (rc) return rc;
    MTG_CUDA_TRY(d2h_chunk(coeffs, B, b0, d_c, C, nb, rec_c, 8, aos, st));
    if (cost) MTG_CUDA_TRY(d2h_chunk(cost, B, b0, d_cost, C, nb, 1, 8, true, st));
    if (free_constraints && rec_free)
      MTG_CUDA_TRY(d2h_chunk(free_constraints, B, b0, d_free, C, nb, rec_free, 8, aos, st));
    if (status) MTG_CUDA_TRY(d2h_chunk(status, B, b0, d_status, C, nb, 1, 4, true, st));
  }
  for (int s = 0; s < slots; ++s) MTG_CUDA_TRY(cudaStreamSynchronize(ctx->stage_stream[s]));
  return MTG_OK;
}

}  // extern "C"

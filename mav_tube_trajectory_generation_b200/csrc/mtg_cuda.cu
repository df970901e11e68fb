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
#include "eval.cuh"

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
  DeviceBuffer stage[kStageSlots];
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

// One batched tensor of a host-memory call.
struct HostTensor {
  const void* host;  // caller pointer (may be null = absent)
  size_t rec;        // elements per trajectory record
  size_t elem;       // bytes per element
  bool input;        // copied H2D before the launch, else D2H after it
  bool vector;       // a plain [B] vector: contiguous in both layouts
  void* dev;         // chunk-local device pointer handed to the launcher
};

// Host-memory mode shared by all entry points: splits the batch into chunks,
// round-robins them over kStageSlots streams (H2D -> kernels -> D2H per chunk, so
// that copies of one chunk overlap the kernels and copies of its neighbours) and
// returns when every output is in place. launch(nb, C, stream) reads ts[i].dev.
template <class Launch>
int run_chunked(mtg_ctx* ctx, cudaStream_t user_stream, size_t B, bool aos, std::vector<HostTensor>& ts,
                Launch&& launch) {
  size_t bytes_per_traj = 0;
  for (auto& t : ts)
    if (t.host) bytes_per_traj += t.rec * t.elem;
  if (bytes_per_traj == 0 || B == 0) return MTG_OK;
  size_t C = 8192;
  if (const char* env = std::getenv("MTG_HOST_CHUNK")) C = std::max(1, std::atoi(env));
  const size_t budget = (size_t)384 << 20;  // per staging slot
  C = std::max<size_t>(1, std::min(C, budget / bytes_per_traj));
  C = std::min(C, B);
  size_t slot_bytes = 0;
  for (auto& t : ts)
    if (t.host) slot_bytes += align256(t.rec * t.elem * C);
  const int n_chunks = (int)((B + C - 1) / C);
  const int slots = std::min(kStageSlots, n_chunks);
  for (int s = 0; s < slots; ++s)
    if (ctx->stage[s].ensure(slot_bytes)) return fail(ctx, MTG_ERR_CUDA, "cudaMalloc of staging buffers failed");
  MTG_CUDA_TRY(cudaStreamSynchronize(user_stream));  // order after work queued on the caller's stream
  for (int c = 0; c < n_chunks; ++c) {
    const int s = c % kStageSlots;
    cudaStream_t st = ctx->stage_stream[s];
    const size_t b0 = (size_t)c * C, nb = std::min(C, B - b0);
    char* base = (char*)ctx->stage[s].ptr;
    size_t off = 0;
    for (auto& t : ts) {
      t.dev = nullptr;
      if (!t.host) continue;
      t.dev = base + off;
      off += align256(t.rec * t.elem * C);
      if (t.input) MTG_CUDA_TRY(h2d_chunk(t.dev, C, t.host, B, b0, nb, t.rec, t.elem, aos || t.vector, st));
    }
    const int rc = launch((int)nb, (int)C, st);
    if (rc) return rc;
    for (auto& t : ts)
      if (t.host && !t.input)
        MTG_CUDA_TRY(d2h_chunk(const_cast<void*>(t.host), B, b0, t.dev, C, nb, t.rec, t.elem, aos || t.vector, st));
  }
  for (int s = 0; s < slots; ++s) MTG_CUDA_TRY(cudaStreamSynchronize(ctx->stage_stream[s]));
  return MTG_OK;
}

template <int NT, bool AOS>
int launch_eval_range_d(mtg_ctx* ctx, int D, const mtg::EvalParams& p, cudaStream_t s) {
  const int block = 128, grid = (p.nb + block - 1) / block;
  if (grid == 0) return MTG_OK;
  switch (D) {
    case 1: mtg::eval_range_kernel<NT, 1, AOS><<<grid, block, 0, s>>>(p); break;
    case 2: mtg::eval_range_kernel<NT, 2, AOS><<<grid, block, 0, s>>>(p); break;
    case 3: mtg::eval_range_kernel<NT, 3, AOS><<<grid, block, 0, s>>>(p); break;
    case 4: mtg::eval_range_kernel<NT, 4, AOS><<<grid, block, 0, s>>>(p); break;
    default: return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "D must be 1..4");
  }
  ++ctx->launches;
  MTG_CUDA_TRY(cudaGetLastError());
  return MTG_OK;
}
int launch_eval_range(mtg_ctx* ctx, int D, bool aos, const mtg::EvalParams& p, cudaStream_t s) {
  if (p.N == 10) return aos ? launch_eval_range_d<10, true>(ctx, D, p, s) : launch_eval_range_d<10, false>(ctx, D, p, s);
  return aos ? launch_eval_range_d<12, true>(ctx, D, p, s) : launch_eval_range_d<12, false>(ctx, D, p, s);
}

template <int NT, bool AOS>
int launch_feasibility_d(mtg_ctx* ctx, int D, const mtg::EvalParams& p, cudaStream_t s) {
  const int block = 128, grid = (p.nb + block - 1) / block;
  if (grid == 0) return MTG_OK;
  switch (D) {
    case 1: mtg::feasibility_kernel<NT, 1, AOS><<<grid, block, 0, s>>>(p); break;
    case 2: mtg::feasibility_kernel<NT, 2, AOS><<<grid, block, 0, s>>>(p); break;
    case 3: mtg::feasibility_kernel<NT, 3, AOS><<<grid, block, 0, s>>>(p); break;
    case 4: mtg::feasibility_kernel<NT, 4, AOS><<<grid, block, 0, s>>>(p); break;
    default: return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "D must be 1..4");
  }
  ++ctx->launches;
  MTG_CUDA_TRY(cudaGetLastError());
  return MTG_OK;
}
int launch_feasibility(mtg_ctx* ctx, int D, bool aos, const mtg::EvalParams& p, cudaStream_t s) {
  if (p.N == 10) return aos ? launch_feasibility_d<10, true>(ctx, D, p, s) : launch_feasibility_d<10, false>(ctx, D, p, s);
  return aos ? launch_feasibility_d<12, true>(ctx, D, p, s) : launch_feasibility_d<12, false>(ctx, D, p, s);
}

int validate_eval(mtg_ctx* ctx, const mtg_problem_desc* desc, int derivative, int max_samples) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (derivative < 0) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "derivative must be >= 0");
  if (max_samples < 0) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "max_samples must be >= 0");
  return MTG_OK;
}

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
    ctx->stage[i].release();
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

/* ------------------------------------------------------------ evaluation */
int mtg_max_time_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* seg_times, double* max_time,
                       void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (!seg_times || !max_time) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "seg_times and max_time are required");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  const int K = desc->K;
  auto launch = [&](const double* t, double* out, int Bld, int nb, cudaStream_t st) {
    const int block = 256, grid = (nb + block - 1) / block;
    if (aos)
      mtg::max_time_kernel<true><<<grid, block, 0, st>>>(t, out, Bld, 0, nb, K);
    else
      mtg::max_time_kernel<false><<<grid, block, 0, st>>>(t, out, Bld, 0, nb, K);
    ++ctx->launches;
    return cudaGetLastError() == cudaSuccess ? MTG_OK : fail(ctx, MTG_ERR_CUDA, "max_time_kernel launch failed");
  };
  if (desc->memory == MTG_MEM_DEVICE) return launch(seg_times, max_time, desc->B, desc->B, stream);
  std::vector<HostTensor> ts = {{seg_times, (size_t)K, 8, true, false, nullptr}, {max_time, 1, 8, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)desc->B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    return launch((const double*)ts[0].dev, (double*)ts[1].dev, C, nb, st);
  });
}

int mtg_eval_range_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                         const double* seg_times, const double* t_start, const double* t_end,
                         const double* dt, int derivative, int max_samples, double* samples,
                         double* sampling_times, int32_t* segment_idx, int32_t* n_samples,
                         uint32_t* status, void* stream_) {
  int rc = validate_eval(ctx, desc, derivative, max_samples);
  if (rc) return rc;
  if (!coeffs || !seg_times || !t_start || !t_end || !dt)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "coeffs, seg_times, t_start, t_end and dt are required");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  rc = ensure_tables(ctx, desc->N, desc->derivative_to_optimize);
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N;
  mtg::EvalParams p = {};
  p.K = K;
  p.N = N;
  p.derivative = derivative;
  p.max_samples = max_samples;
  if (desc->memory == MTG_MEM_DEVICE) {
    p.coeffs = coeffs; p.seg_times = seg_times; p.t_start = t_start; p.t_end = t_end; p.dt = dt;
    p.samples = samples; p.sampling_times = sampling_times; p.segment_idx = segment_idx;
    p.n_samples = n_samples; p.status = status;
    p.B = B; p.b0 = 0; p.nb = B;
    p.vec_ok = ((uintptr_t)coeffs % 16 == 0) ? 1 : 0;
    return launch_eval_range(ctx, D, aos, p, stream);
  }
  const size_t S = (size_t)max_samples;
  std::vector<HostTensor> ts = {
      {coeffs, (size_t)K * D * N, 8, true, false, nullptr}, {seg_times, (size_t)K, 8, true, false, nullptr},
      {t_start, 1, 8, true, true, nullptr},  {t_end, 1, 8, true, true, nullptr},
      {dt, 1, 8, true, true, nullptr},       {S ? samples : nullptr, S * D, 8, false, false, nullptr},
      {S ? sampling_times : nullptr, S, 8, false, false, nullptr},
      {S ? segment_idx : nullptr, S, 4, false, false, nullptr},
      {n_samples, 1, 4, false, true, nullptr}, {status, 1, 4, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    p.coeffs = (const double*)ts[0].dev; p.seg_times = (const double*)ts[1].dev;
    p.t_start = (const double*)ts[2].dev; p.t_end = (const double*)ts[3].dev; p.dt = (const double*)ts[4].dev;
    p.samples = (double*)ts[5].dev; p.sampling_times = (double*)ts[6].dev; p.segment_idx = (int32_t*)ts[7].dev;
    p.n_samples = (int32_t*)ts[8].dev; p.status = (uint32_t*)ts[9].dev;
    p.B = C; p.b0 = 0; p.nb = nb; p.vec_ok = 1;
    return launch_eval_range(ctx, D, aos, p, st);
  });
}

int mtg_eval_at_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                      const double* seg_times, const double* t, int M, int derivative, double* out,
                      int32_t* segment_idx, uint32_t* status, void* stream_) {
  int rc = validate_eval(ctx, desc, derivative, M);
  if (rc) return rc;
  if (!coeffs || !seg_times || !t || !out)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "coeffs, seg_times, t and out are required");
  if (desc->B == 0 || M == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  rc = ensure_tables(ctx, desc->N, desc->derivative_to_optimize);
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N;
  mtg::EvalAtParams p = {};
  p.K = K; p.N = N; p.D = D; p.M = M; p.derivative = derivative;
  auto launch = [&](cudaStream_t st) {
    if (p.status) {
      cudaError_t e = cudaMemsetAsync(p.status, 0, sizeof(uint32_t) * p.nb, st);
      if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMemsetAsync(status)");
    }
    const long long total = (long long)p.nb * M;
    const int block = 256;
    const long long grid = (total + block - 1) / block;
    if (aos)
      mtg::eval_at_kernel<true><<<(unsigned)grid, block, 0, st>>>(p);
    else
      mtg::eval_at_kernel<false><<<(unsigned)grid, block, 0, st>>>(p);
    ++ctx->launches;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? MTG_OK : cuda_fail(ctx, e, "eval_at_kernel");
  };
  if (desc->memory == MTG_MEM_DEVICE) {
    p.coeffs = coeffs; p.seg_times = seg_times; p.t = t; p.out = out; p.segment_idx = segment_idx;
    p.status = status; p.B = B; p.b0 = 0; p.nb = B;
    return launch(stream);
  }
  std::vector<HostTensor> ts = {
      {coeffs, (size_t)K * D * N, 8, true, false, nullptr}, {seg_times, (size_t)K, 8, true, false, nullptr},
      {t, (size_t)M, 8, true, false, nullptr},              {out, (size_t)M * D, 8, false, false, nullptr},
      {segment_idx, (size_t)M, 4, false, false, nullptr},   {status, 1, 4, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    p.coeffs = (const double*)ts[0].dev; p.seg_times = (const double*)ts[1].dev; p.t = (const double*)ts[2].dev;
    p.out = (double*)ts[3].dev; p.segment_idx = (int32_t*)ts[4].dev; p.status = (uint32_t*)ts[5].dev;
    p.B = C; p.b0 = 0; p.nb = nb;
    return launch(st);
  });
}

int mtg_feasibility_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                          const double* seg_times, const double* positions, const double* radii,
                          double v_max, double a_max, const double* t_start, const double* t_end,
                          const double* dt, int max_samples, double* samples, uint8_t* flags,
                          double* max_v, double* max_a, uint8_t* feasible, int32_t* n_samples,
                          uint32_t* status, void* stream_) {
  int rc = validate_eval(ctx, desc, 0, max_samples);
  if (rc) return rc;
  if (!coeffs || !seg_times || !t_start || !t_end || !dt)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "coeffs, seg_times, t_start, t_end and dt are required");
  if (radii && (desc->D != 3 || !positions))
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "the tube check needs D == 3 and the vertex positions");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  rc = ensure_tables(ctx, desc->N, desc->derivative_to_optimize);
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N;
  mtg::EvalParams p = {};
  p.K = K; p.N = N; p.derivative = 0; p.max_samples = max_samples; p.v_max = v_max; p.a_max = a_max;
  if (desc->memory == MTG_MEM_DEVICE) {
    p.coeffs = coeffs; p.seg_times = seg_times; p.t_start = t_start; p.t_end = t_end; p.dt = dt;
    p.samples = samples; p.flags = flags; p.max_v = max_v; p.max_a = max_a; p.feasible = feasible;
    p.positions = radii ? positions : nullptr; p.radii = radii;
    p.n_samples = n_samples; p.status = status;
    p.B = B; p.b0 = 0; p.nb = B;
    p.vec_ok = ((uintptr_t)coeffs % 16 == 0) ? 1 : 0;
    return launch_feasibility(ctx, D, aos, p, stream);
  }
  const size_t S = (size_t)max_samples;
  std::vector<HostTensor> ts = {
      {coeffs, (size_t)K * D * N, 8, true, false, nullptr}, {seg_times, (size_t)K, 8, true, false, nullptr},
      {t_start, 1, 8, true, true, nullptr}, {t_end, 1, 8, true, true, nullptr}, {dt, 1, 8, true, true, nullptr},
      {radii ? positions : nullptr, (size_t)(K + 1) * 3, 8, true, false, nullptr},
      {radii, (size_t)K * 2, 8, true, false, nullptr},
      {S ? samples : nullptr, S * D, 8, false, false, nullptr}, {S ? flags : nullptr, S, 1, false, false, nullptr},
      {max_v, 1, 8, false, true, nullptr}, {max_a, 1, 8, false, true, nullptr},
      {feasible, 1, 1, false, true, nullptr}, {n_samples, 1, 4, false, true, nullptr},
      {status, 1, 4, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    p.coeffs = (const double*)ts[0].dev; p.seg_times = (const double*)ts[1].dev;
    p.t_start = (const double*)ts[2].dev; p.t_end = (const double*)ts[3].dev; p.dt = (const double*)ts[4].dev;
    p.positions = (const double*)ts[5].dev; p.radii = (const double*)ts[6].dev;
    p.samples = (double*)ts[7].dev; p.flags = (uint8_t*)ts[8].dev;
    p.max_v = (double*)ts[9].dev; p.max_a = (double*)ts[10].dev; p.feasible = (uint8_t*)ts[11].dev;
    p.n_samples = (int32_t*)ts[12].dev; p.status = (uint32_t*)ts[13].dev;
    p.B = C; p.b0 = 0; p.nb = nb; p.vec_ok = 1;
    return launch_feasibility(ctx, D, aos, p, st);
  });
}

}  // extern "C"

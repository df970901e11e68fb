// Host-side plumbing shared by the translation units of libmtg_cuda.so: the
// context, error helpers, argument validation and the chunked host-memory mode.
// Each kernel family lives in its own .cu (compiled in parallel by _build.py);
// each has a private copy of the constant tables, registered here.
#ifndef MTG_HOST_COMMON_H_
#define MTG_HOST_COMMON_H_

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mtg_cuda.h"
#include "device_tables.cuh"

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

struct mtg_ctx {
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  std::string err;
  uint64_t launches = 0;
  int solve_overlap = 0;  // mtg_set_solve_overlap
  cudaStream_t stage_stream[kStageSlots] = {nullptr, nullptr, nullptr};
  DeviceBuffer stage[kStageSlots];  // one staging arena per slot (host-memory mode)
  DeviceBuffer scratch;             // small per-context device scratch
  // per-stream scratch (tube constants of the sweep): launches on different streams never share one
  std::vector<std::pair<cudaStream_t, DeviceBuffer>> stream_scratch;
  DeviceBuffer* scratch_for(cudaStream_t s) {
    for (auto& e : stream_scratch)
      if (e.first == s) return &e.second;
    stream_scratch.emplace_back(s, DeviceBuffer());
    return &stream_scratch.back().second;
  }
  // per-stream argmin state: block partials + the ticket counter of the single-launch reduction (zeroed once;
  // the kernel leaves it at zero). Never shared with the scratch above, which other kernels overwrite.
  std::vector<std::pair<cudaStream_t, DeviceBuffer>> stream_argmin;
  DeviceBuffer* argmin_state_for(cudaStream_t s) {
    for (auto& e : stream_argmin)
      if (e.first == s) return &e.second;
    stream_argmin.emplace_back(s, DeviceBuffer());
    return &stream_argmin.back().second;
  }
  // per-stream work space of the non-linear objective (perturbed segments, nominal maxima, gradients)
  std::vector<std::pair<std::pair<cudaStream_t, int>, DeviceBuffer>> stream_nl;
  DeviceBuffer* nl_scratch_for(cudaStream_t s, int slot) {
    for (auto& e : stream_nl)
      if (e.first.first == s && e.first.second == slot) return &e.second;
    stream_nl.emplace_back(std::make_pair(s, slot), DeviceBuffer());
    return &stream_nl.back().second;
  }
  void* nccl = nullptr;             // lazily created NCCL state (argmin gather)
  cudaEvent_t order_event = nullptr;  // host-memory mode: orders the staging streams behind the caller's stream
};

namespace mtg {

inline int fail(mtg_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}
inline int cuda_fail(mtg_ctx* ctx, cudaError_t e, const char* where) {
  return fail(ctx, MTG_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}

#define MTG_CUDA_TRY(expr)                                        \
  do {                                                            \
    cudaError_t _e = (expr);                                      \
    if (_e != cudaSuccess) return mtg::cuda_fail(ctx, _e, #expr); \
  } while (0)

// core.cu
int validate_desc(mtg_ctx* ctx, const mtg_problem_desc* d);
// The per-(N, derivative) constant tables are ONE mutable set per device. TableGuard makes a launch
// sequence safe against that: it takes the library-wide table lock, makes (N, derivative) resident
// (a switch drains the device first, so no in-flight kernel sees a torn table, and synchronises again
// after the upload, so no later kernel on any stream can start before the copy has landed) and keeps
// the lock until it is destroyed, i.e. until every kernel of the call has been ENQUEUED. Two host
// threads using different (N, derivative) on one device therefore serialise at the switch: a kernel
// enqueued under one table set has finished before the other set is uploaded. A switch is refused while
// `stream` is being captured (cudaDeviceSynchronize would invalidate the capture): run one call with
// that (N, derivative) before capturing.
class TableGuard {
 public:
  TableGuard(mtg_ctx* ctx, int N, int derivative, cudaStream_t stream);
  ~TableGuard();
  TableGuard(const TableGuard&) = delete;
  TableGuard& operator=(const TableGuard&) = delete;
  int rc() const { return rc_; }

 private:
  int rc_;
  bool locked_;
};
// uploads the (N, derivative)-independent base table to this device once (mtg_create)
int ensure_base(mtg_ctx* ctx);
// every kernel translation unit registers the uploaders of its private c_tab / c_base copies
typedef cudaError_t (*TableUploader)(const DevTables*);
typedef cudaError_t (*BaseUploader)(const DevBase*);
void register_table_uploader(TableUploader f);
void register_base_uploader(BaseUploader f);
#define MTG_REGISTER_TABLES()                                                               \
  namespace {                                                                               \
  cudaError_t upload_tables_(const mtg::DevTables* h) {                                     \
    return cudaMemcpyToSymbol(mtg::c_tab, h, sizeof(mtg::DevTables));                       \
  }                                                                                         \
  struct TableRegistrar_ {                                                                  \
    TableRegistrar_() { mtg::register_table_uploader(&upload_tables_); }                    \
  } table_registrar_;                                                                       \
  }
#define MTG_REGISTER_BASE()                                                                 \
  namespace {                                                                               \
  cudaError_t upload_base_(const mtg::DevBase* h) {                                         \
    return cudaMemcpyToSymbol(mtg::c_base, h, sizeof(mtg::DevBase));                        \
  }                                                                                         \
  struct BaseRegistrar_ {                                                                   \
    BaseRegistrar_() { mtg::register_base_uploader(&upload_base_); }                        \
  } base_registrar_;                                                                        \
  }

// eval_tm.cu: time-major sweep (trajectory-contiguous sample outputs, AoS layout)
struct EvalParams;
bool eval_tm_supported(const EvalParams& p);
int launch_eval_tm(mtg_ctx* ctx, int D, bool feasibility, const EvalParams& p, cudaStream_t s);

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Copies of a chunk [b0, b0+nb) of a batched tensor with `rec` elements per record
// between the caller's host tensor (batch B) and a chunk-sized device tensor (batch C).
inline cudaError_t h2d_chunk(void* dst, size_t C, const void* src, size_t B, size_t b0, size_t nb, size_t rec,
                             size_t elem, bool aos, cudaStream_t s) {
  if (aos)
    return cudaMemcpyAsync(dst, (const char*)src + b0 * rec * elem, nb * rec * elem, cudaMemcpyHostToDevice, s);
  return cudaMemcpy2DAsync(dst, C * elem, (const char*)src + b0 * elem, B * elem, nb * elem, rec,
                           cudaMemcpyHostToDevice, s);
}
inline cudaError_t d2h_chunk(void* dst, size_t B, size_t b0, const void* src, size_t C, size_t nb, size_t rec,
                             size_t elem, bool aos, cudaStream_t s) {
  if (aos)
    return cudaMemcpyAsync((char*)dst + b0 * rec * elem, src, nb * rec * elem, cudaMemcpyDeviceToHost, s);
  return cudaMemcpy2DAsync((char*)dst + b0 * elem, B * elem, src, C * elem, nb * elem, rec,
                           cudaMemcpyDeviceToHost, s);
}

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
  // chunk: 8,192 trajectories, more when a trajectory moves few bytes (a cost-only solve: 356 B) so that a
  // chunk is still ~6 MB of PCIe traffic, but never fewer than 4 chunks per call (measured: +10 % on cost-only)
  size_t C = 8192;
  C = std::max(C, std::min((((size_t)6 << 20) / bytes_per_traj) & ~(size_t)1023, (B / 4) & ~(size_t)1023));
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
  // order the staging streams after work already queued on the caller's stream (no host wait)
  if (!ctx->order_event) MTG_CUDA_TRY(cudaEventCreateWithFlags(&ctx->order_event, cudaEventDisableTiming));
  MTG_CUDA_TRY(cudaEventRecord(ctx->order_event, user_stream));
  for (int s = 0; s < slots; ++s) MTG_CUDA_TRY(cudaStreamWaitEvent(ctx->stage_stream[s], ctx->order_event, 0));
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

}  // namespace mtg
#endif

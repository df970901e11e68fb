// libmtg_cuda.so — context management, constant tables and argument validation.
#include <mutex>

#include "host_common.h"

namespace mtg {

namespace {
std::recursive_mutex g_tab_mutex;
int g_tab_N[64];
int g_tab_d[64];
bool g_base_done[64];
bool g_tab_init = false;
std::vector<TableUploader>& uploaders() {
  static std::vector<TableUploader> v;
  return v;
}
std::vector<BaseUploader>& base_uploaders() {
  static std::vector<BaseUploader> v;
  return v;
}
void init_state() {
  if (g_tab_init) return;
  for (int i = 0; i < 64; ++i) {
    g_tab_N[i] = g_tab_d[i] = -1;
    g_base_done[i] = false;
  }
  g_tab_init = true;
}
}  // namespace

void register_table_uploader(TableUploader f) { uploaders().push_back(f); }
void register_base_uploader(BaseUploader f) { base_uploaders().push_back(f); }

int ensure_base(mtg_ctx* ctx) {
  std::lock_guard<std::recursive_mutex> lock(g_tab_mutex);
  init_state();
  const int dev = ctx->device & 63;
  if (g_base_done[dev]) return MTG_OK;
  Tables t;
  if (!compute_tables(2, 0, &t)) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "base table");
  DevBase h;
  std::memcpy(h.base, t.base, sizeof(h.base));
  for (BaseUploader f : base_uploaders()) MTG_CUDA_TRY(f(&h));
  MTG_CUDA_TRY(cudaDeviceSynchronize());  // the copies have landed before any kernel of any stream can start
  g_base_done[dev] = true;
  return MTG_OK;
}

namespace {
// Makes (N, derivative) the resident table set of this device in every kernel translation unit.
// Called with g_tab_mutex held.
int switch_tables(mtg_ctx* ctx, int N, int derivative, cudaStream_t stream) {
  init_state();
  const int dev = ctx->device & 63;
  if (g_tab_N[dev] == N && g_tab_d[dev] == derivative) return MTG_OK;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone)
    return fail(ctx, MTG_ERR_UNSUPPORTED,
                "the (N, derivative) tables must be resident before stream capture: run one call with this "
                "(N, derivative_to_optimize) outside the capture first");
  Tables t;
  if (!compute_tables(N, derivative, &t))
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "invalid (N, derivative_to_optimize)");
  DevTables h;
  std::memcpy(h.H1, t.H1, sizeof(h.H1));
  std::memcpy(h.Ainv1, t.Ainv1, sizeof(h.Ainv1));
  std::memcpy(h.W, t.W, sizeof(h.W));
  std::memcpy(h.Lt, t.Lt, sizeof(h.Lt));
  for (int j = 0; j < MTG_TAB_LD; ++j) h.inv_factorial[j] = 1.0 / t.base[j * MTG_BASE_LD + j];
  h.N = N;
  h.derivative = derivative;
  MTG_CUDA_TRY(cudaDeviceSynchronize());  // drain: no in-flight kernel sees a torn table
  g_tab_N[dev] = g_tab_d[dev] = -1;
  for (TableUploader f : uploaders()) MTG_CUDA_TRY(f(&h));
  MTG_CUDA_TRY(cudaDeviceSynchronize());  // landed: non-blocking streams are not ordered behind the copy
  g_tab_N[dev] = N;
  g_tab_d[dev] = derivative;
  return MTG_OK;
}
}  // namespace

TableGuard::TableGuard(mtg_ctx* ctx, int N, int derivative, cudaStream_t stream) : rc_(MTG_OK), locked_(true) {
  g_tab_mutex.lock();
  rc_ = switch_tables(ctx, N, derivative, stream);
}
TableGuard::~TableGuard() {
  if (locked_) g_tab_mutex.unlock();
}

void destroy_nccl_state(mtg_ctx* ctx);  // argmin.cu

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

}  // namespace mtg

using namespace mtg;

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
  if (mtg::ensure_base(ctx) != MTG_OK) {
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
  mtg::destroy_nccl_state(ctx);
  for (int i = 0; i < kStageSlots; ++i) {
    ctx->stage[i].release();
    if (ctx->stage_stream[i]) cudaStreamDestroy(ctx->stage_stream[i]);
  }
  if (ctx->order_event) cudaEventDestroy(ctx->order_event);
  ctx->scratch.release();
  for (auto& e : ctx->stream_scratch) e.second.release();
  for (auto& e : ctx->stream_argmin) e.second.release();
  for (auto& e : ctx->stream_nl) e.second.release();
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

}  // extern "C"

// libmtg_cuda.so — candidate-sweep reduction: local argmin of computeCost() over a batch and the
// one collective of the sharded sweep: an all-gather of one 16-byte
// {cost, index} pair per rank over NCCL (NVLink 5 / NVSwitch), after which every rank holds the
// same global argmin. NCCL is bound at run time (dlopen): the library has no link-time
// dependency on it and single-GPU users never load it.
#include <dlfcn.h>

#include <cmath>
#include <limits>

#include "argmin.cuh"
#include "host_common.h"

using namespace mtg;

namespace {

constexpr int kArgminBlock = 256;

// block-wide reduction of (c, i); the result is valid in thread 0
__device__ __forceinline__ void block_reduce(double& c, long long& i, double* sc, long long* si) {
  warp_reduce(c, i);
  if ((threadIdx.x & 31) == 0) {
    sc[threadIdx.x >> 5] = c;
    si[threadIdx.x >> 5] = i;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    c = threadIdx.x < kArgminBlock / 32 ? sc[threadIdx.x] : INFINITY;
    i = threadIdx.x < kArgminBlock / 32 ? si[threadIdx.x] : kInfIdx;
    warp_reduce(c, i);
  }
  __syncthreads();
}

// ONE launch: every block scans its grid-stride share and publishes one pair; the block that draws the
// last ticket folds the pairs (+ the running best) into out. atomicInc wraps the ticket back to zero, so
// the counter needs no reset between launches.
__global__ void __launch_bounds__(kArgminBlock) argmin_kernel(const double* __restrict__ cost,
                                                              const uint32_t* __restrict__ status, long long n,
                                                              long long offset, Best* partial, unsigned* ticket,
                                                              const Best* __restrict__ running, Best* out) {
  __shared__ double sc[kArgminBlock / 32];
  __shared__ long long si[kArgminBlock / 32];
  __shared__ bool last;
  double c = INFINITY;
  long long i = kInfIdx;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
    const double v = cost[q];
    if (status && status[q] != 0u) continue;  // failed solves never win
    if (!(v == v)) continue;                  // NaN
    if (better(v, offset + q, c, i)) {
      c = v;
      i = offset + q;
    }
  }
  block_reduce(c, i, sc, si);
  if (threadIdx.x == 0) {
    Best b;
    b.cost = c;
    b.idx = i;
    partial[blockIdx.x] = b;
    __threadfence();
    last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  c = INFINITY;
  i = kInfIdx;
  for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) {
    const double pc = __ldcg(&partial[q].cost);  // written by other blocks: read through L2
    const long long pi = __ldcg(&partial[q].idx);
    if (better(pc, pi, c, i)) {
      c = pc;
      i = pi;
    }
  }
  if (running && threadIdx.x == 0) {
    const Best b = *running;
    if (b.idx >= 0 && better(b.cost, b.idx, c, i)) {
      c = b.cost;
      i = b.idx;
    }
  }
  block_reduce(c, i, sc, si);
  if (threadIdx.x == 0) {
    Best b;
    b.cost = c;
    b.idx = (i == kInfIdx) ? -1 : i;
    *out = b;
  }
}

// Final selection of the sharded sweep ON THE DEVICE: folds the `world` gathered pairs into one (same
// total order as everywhere else). One warp; the pairs were written by NCCL on this stream.
__global__ void fold_pairs_kernel(const Best* __restrict__ pairs, int world, Best* out) {
  double c = INFINITY;
  long long i = kInfIdx;
  for (int q = threadIdx.x; q < world; q += 32) {
    const Best b = pairs[q];
    if (b.idx >= 0 && b.idx != kInfIdx && b.cost == b.cost && better(b.cost, b.idx, c, i)) {
      c = b.cost;
      i = b.idx;
    }
  }
  warp_reduce(c, i);
  if (threadIdx.x == 0) {
    Best b;
    b.cost = c;
    b.idx = (i == kInfIdx) ? -1 : i;
    *out = b;
  }
}

// ---------------------------------------------------------------- NCCL, bound at run time
typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  Best* gathered = nullptr;  // device [world]
};
constexpr int kNcclChar = 0;  // ncclInt8 / ncclChar

int load_nccl(mtg_ctx* ctx, NcclApi** out) {
  if (ctx->nccl) {
    *out = (NcclApi*)ctx->nccl;
    return MTG_OK;
  }
  const char* names[] = {std::getenv("MTG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* nme : names) {
    if (!nme || !*nme) continue;
    h = dlopen(nme, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail(ctx, MTG_ERR_NCCL, std::string("cannot load libnccl.so.2 (set MTG_NCCL_LIB): ") + dlerror());
  NcclApi* api = new NcclApi();
  api->handle = h;
  api->GetUniqueId = (int (*)(ncclUniqueId*))dlsym(h, "ncclGetUniqueId");
  api->CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
  api->CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
  api->AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllGather");
  api->GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!api->GetUniqueId || !api->CommInitRank || !api->CommDestroy || !api->AllGather) {
    delete api;
    return fail(ctx, MTG_ERR_NCCL, "libnccl is missing a required symbol");
  }
  ctx->nccl = api;
  *out = api;
  return MTG_OK;
}

int nccl_fail(mtg_ctx* ctx, NcclApi* api, int rc, const char* what) {
  return fail(ctx, MTG_ERR_NCCL,
              std::string(what) + ": " + (api->GetErrorString ? api->GetErrorString(rc) : "NCCL error"));
}

int local_argmin(mtg_ctx* ctx, const double* cost, const uint32_t* status, long long n, long long offset,
                 int accumulate, Best* best_dev, cudaStream_t s) {
  const int max_blocks = 4 * std::max(ctx->sm_count, 1);
  const int blocks = (int)std::max<long long>(1, std::min<long long>(max_blocks, (n + kArgminBlock - 1) / kArgminBlock));
  DeviceBuffer* state = ctx->argmin_state_for(s);
  if (!state->ptr) {
    if (state->ensure(256 + (size_t)max_blocks * sizeof(Best))) return fail(ctx, MTG_ERR_CUDA, "cudaMalloc of the argmin state failed");
    MTG_CUDA_TRY(cudaMemset(state->ptr, 0, 256));  // the ticket counter (first 4 bytes), once
  }
  unsigned* ticket = (unsigned*)state->ptr;
  Best* partial = (Best*)((char*)state->ptr + 256);
  argmin_kernel<<<blocks, kArgminBlock, 0, s>>>(cost, status, n, offset, partial, ticket, accumulate ? best_dev : nullptr,
                                               best_dev);
  ctx->launches += 1;
  MTG_CUDA_TRY(cudaGetLastError());
  return MTG_OK;
}

}  // namespace

namespace mtg {
void destroy_nccl_state(mtg_ctx* ctx) {
  if (!ctx->nccl) return;
  NcclApi* api = (NcclApi*)ctx->nccl;
  if (api->comm) api->CommDestroy(api->comm);
  if (api->gathered) cudaFree(api->gathered);
  delete api;  // the library handle stays loaded for the life of the process
  ctx->nccl = nullptr;
}
}  // namespace mtg

extern "C" {

int mtg_argmin_batch(mtg_ctx* ctx, const double* cost, const uint32_t* status, int64_t n, int64_t global_offset,
                     int accumulate, void* best, void* stream_) {
  if (!ctx) return MTG_ERR_INVALID_ARGUMENT;
  if (!best || (n > 0 && !cost) || n < 0) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "cost and best are required");
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  return local_argmin(ctx, cost, status, n, global_offset, accumulate, (Best*)best, (cudaStream_t)stream_);
}

int mtg_nccl_unique_id(mtg_ctx* ctx, uint8_t id[128]) {
  if (!ctx || !id) return MTG_ERR_INVALID_ARGUMENT;
  NcclApi* api;
  int rc = load_nccl(ctx, &api);
  if (rc) return rc;
  ncclUniqueId u;
  const int nrc = api->GetUniqueId(&u);
  if (nrc) return nccl_fail(ctx, api, nrc, "ncclGetUniqueId");
  std::memcpy(id, u.internal, 128);
  return MTG_OK;
}

int mtg_nccl_init(mtg_ctx* ctx, const uint8_t id[128], int rank, int world) {
  if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return MTG_ERR_INVALID_ARGUMENT;
  NcclApi* api;
  int rc = load_nccl(ctx, &api);
  if (rc) return rc;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  if (api->comm) {
    api->CommDestroy(api->comm);
    api->comm = nullptr;
  }
  ncclUniqueId u;
  std::memcpy(u.internal, id, 128);
  const int nrc = api->CommInitRank(&api->comm, world, u, rank);
  if (nrc) return nccl_fail(ctx, api, nrc, "ncclCommInitRank");
  api->rank = rank;
  api->world = world;
  if (api->gathered) cudaFree(api->gathered);
  MTG_CUDA_TRY(cudaMalloc((void**)&api->gathered, sizeof(Best) * (size_t)(world + 1)));
  return MTG_OK;
}

int mtg_best_allgather(mtg_ctx* ctx, const void* best_local, void* best_global, void* stream_) {
  if (!ctx) return MTG_ERR_INVALID_ARGUMENT;
  if (!best_local || !best_global) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "best_local and best_global are required");
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream_;
  NcclApi* api = (NcclApi*)ctx->nccl;
  const int world = (api && api->comm) ? api->world : 1;
  const Best* pairs = (const Best*)best_local;
  if (world > 1) {
    const int nrc = api->AllGather(best_local, api->gathered, sizeof(Best), kNcclChar, api->comm, s);
    if (nrc) return nccl_fail(ctx, api, nrc, "ncclAllGather");
    pairs = api->gathered;
  }
  fold_pairs_kernel<<<1, 32, 0, s>>>(pairs, world, (Best*)best_global);
  ctx->launches += 1;
  MTG_CUDA_TRY(cudaGetLastError());
  return MTG_OK;
}

int mtg_argmin_allgather(mtg_ctx* ctx, const double* cost, const uint32_t* status, int64_t n_local,
                         int64_t global_offset, double* best_cost, int64_t* best_idx, void* stream_) {
  if (!ctx) return MTG_ERR_INVALID_ARGUMENT;
  if (!best_cost || !best_idx) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "best_cost and best_idx are required");
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream_;
  NcclApi* api = (NcclApi*)ctx->nccl;
  const int world = (api && api->comm) ? api->world : 1;
  Best host[64];
  if (world > 64) return fail(ctx, MTG_ERR_UNSUPPORTED, "argmin gather supports up to 64 ranks");
  Best* mine;
  if (world == 1) {
    if (ctx->scratch.ensure(sizeof(Best))) return fail(ctx, MTG_ERR_CUDA, "cudaMalloc failed");
    mine = (Best*)ctx->scratch.ptr;
  } else {
    mine = api->gathered + world;  // send slot behind the receive buffer
  }
  int rc = local_argmin(ctx, cost, status, n_local, global_offset, 0, mine, s);
  if (rc) return rc;
  if (world == 1) {
    MTG_CUDA_TRY(cudaMemcpyAsync(host, mine, sizeof(Best), cudaMemcpyDeviceToHost, s));
  } else {
    // 16 bytes per rank: latency-bound, sent as bytes so that the index stays an exact int64
    const int nrc = api->AllGather(mine, api->gathered, sizeof(Best), kNcclChar, api->comm, s);
    if (nrc) return nccl_fail(ctx, api, nrc, "ncclAllGather");
    MTG_CUDA_TRY(cudaMemcpyAsync(host, api->gathered, sizeof(Best) * world, cudaMemcpyDeviceToHost, s));
  }
  MTG_CUDA_TRY(cudaStreamSynchronize(s));
  double bc = INFINITY;
  long long bi = -1;
  for (int r = 0; r < world; ++r) {
    if (host[r].idx < 0 || host[r].idx == 0x7fffffffffffffffLL) continue;
    if (bi < 0 || host[r].cost < bc || (host[r].cost == bc && host[r].idx < bi)) {
      bc = host[r].cost;
      bi = host[r].idx;
    }
  }
  *best_cost = bc;
  *best_idx = bi;
  return MTG_OK;
}

}  // extern "C"

// libmtg_cuda.so — G1 on the device: the candidate generator of a sweep. Random vertices in a box with the
// reference's 0.2 m rejection rule and its Nfabian segment times, produced where they are consumed instead of
// in numpy + a 344-byte-per-candidate H2D copy.
//
// Replaces (reference): createRandomVertices src/vertex.cpp:27-82 (uniform positions per dimension, resampled
// until |pos - last| > 0.2, :65-72) and estimateSegmentTimesNfabian :252-269. The reference draws from
// std::mt19937(seed), a serial generator; a sweep over millions of candidates needs a COUNTER-BASED one:
// Philox4x32-10 keyed by the sweep seed, counter = (candidate index, draw number), so candidate b is the same
// on any rank, in any batch split and on the host (tests reproduce it in numpy bit for bit).
#include "host_common.h"
#include "solve_canonical.cuh"  // at<AOS>()

using namespace mtg;

namespace {

struct GenerateParams {
  double* __restrict__ positions;  // elem v*D + dim, rec (K+1)*D
  double* __restrict__ seg_times;  // elem i, rec K
  unsigned long long seed;
  long long first_index;           // global index of candidate 0 of this batch
  double lo[4], hi[4];
  double v_max, a_max, magic;
  int B, b0, nb, K, D;
};

__device__ __forceinline__ void philox4x32_10(unsigned (&c)[4], unsigned k0, unsigned k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const unsigned n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// 53-bit uniform in [0, 1) from two 32-bit words
__device__ __forceinline__ double u53(unsigned a, unsigned b) {
  return (double)(((unsigned long long)(a >> 5) << 26) | (unsigned long long)(b >> 6)) * (1.0 / 9007199254740992.0);
}

template <bool AOS>
__global__ void __launch_bounds__(128) generate_kernel(const GenerateParams p) {
  const int local = blockIdx.x * blockDim.x + threadIdx.x;
  if (local >= p.nb) return;
  const int b = p.b0 + local;
  const size_t B = (size_t)p.B;
  const int K = p.K, D = p.D;
  const unsigned long long gidx = (unsigned long long)(p.first_index + b);
  const unsigned k0 = (unsigned)p.seed, k1 = (unsigned)(p.seed >> 32);
  unsigned draw = 0;
  double last[4] = {0.0, 0.0, 0.0, 0.0};
  for (int v = 0; v <= K; ++v) {
    double pos[4] = {0.0, 0.0, 0.0, 0.0};
    double dist = 0.0;
    for (;;) {
      // one draw = one Philox block: four 32-bit words -> two uniforms; D <= 4 needs two blocks
      unsigned c[4] = {(unsigned)gidx, (unsigned)(gidx >> 32), draw, 0u};
      philox4x32_10(c, k0, k1);
      unsigned e[4] = {(unsigned)gidx, (unsigned)(gidx >> 32), draw, 1u};
      if (D > 2) philox4x32_10(e, k0, k1);
      ++draw;
      const double u[4] = {u53(c[0], c[1]), u53(c[2], c[3]), u53(e[0], e[1]), u53(e[2], e[3])};
      double s = 0.0;
#pragma unroll
      for (int dim = 0; dim < 4; ++dim)
        if (dim < D) {
          pos[dim] = __dadd_rn(p.lo[dim], __dmul_rn(u[dim], p.hi[dim] - p.lo[dim]));  // no FMA: the host twin multiplies, then adds
          const double dd = pos[dim] - last[dim];
          s = __dadd_rn(s, __dmul_rn(dd, dd));
        }
      dist = sqrt(s);
      if (v == 0 || dist > 0.2) break;  // vertex.cpp:65-72
    }
#pragma unroll
    for (int dim = 0; dim < 4; ++dim)
      if (dim < D) {
        p.positions[at<AOS>((size_t)v * D + dim, (size_t)(K + 1) * D, B, b)] = pos[dim];
        last[dim] = pos[dim];
      }
    if (v >= 1) {
      // estimateSegmentTimesNfabian, vertex.cpp:252-269
      const double t = dist / p.v_max * 2 * (1.0 + p.magic * p.v_max / p.a_max * exp(-dist / p.v_max * 2));
      p.seg_times[at<AOS>((size_t)(v - 1), (size_t)K, B, b)] = t;
    }
  }
}

}  // namespace

extern "C" int mtg_generate_candidates_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, uint64_t seed,
                                             int64_t first_index, const double* pos_min, const double* pos_max,
                                             double v_max, double a_max, double magic_fabian_constant,
                                             double* positions, double* seg_times, void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (desc->memory != MTG_MEM_DEVICE)
    return fail(ctx, MTG_ERR_UNSUPPORTED, "mtg_generate_candidates_batch writes device tensors");
  if (!pos_min || !pos_max || !positions || !seg_times || !(v_max > 0.0) || !(a_max > 0.0))
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "bounds, outputs, v_max > 0 and a_max > 0 are required");
  for (int k = 0; k < desc->D; ++k)
    if (!(pos_max[k] - pos_min[k] > 0.4))  // the rejection rule needs room (reference CHECKs min < max, vertex.cpp:31-34)
      return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "pos_max - pos_min must exceed 0.4 in every dimension");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  GenerateParams p = {};
  p.positions = positions; p.seg_times = seg_times; p.seed = seed; p.first_index = first_index;
  for (int k = 0; k < desc->D; ++k) {
    p.lo[k] = pos_min[k];
    p.hi[k] = pos_max[k];
  }
  p.v_max = v_max; p.a_max = a_max; p.magic = magic_fabian_constant;
  p.B = desc->B; p.b0 = 0; p.nb = desc->B; p.K = desc->K; p.D = desc->D;
  const int grid = (p.nb + 127) / 128;
  if (desc->layout == MTG_LAYOUT_AOS)
    generate_kernel<true><<<grid, 128, 0, (cudaStream_t)stream_>>>(p);
  else
    generate_kernel<false><<<grid, 128, 0, (cudaStream_t)stream_>>>(p);
  ++ctx->launches;
  MTG_CUDA_TRY(cudaGetLastError());
  return MTG_OK;
}

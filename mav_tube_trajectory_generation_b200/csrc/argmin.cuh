// Device helpers of the candidate-sweep reduction ({cost, global index} pairs), shared by argmin.cu and the
// fused epilogue of solve_canonical_kernel.
#ifndef MTG_ARGMIN_CUH_
#define MTG_ARGMIN_CUH_

#include <math.h>

namespace mtg {

struct Best {
  double cost;
  long long idx;
};

constexpr long long kInfIdx = 0x7fffffffffffffffLL;  // "nothing yet": ordered last

__device__ __forceinline__ bool better(double c, long long i, double bc, long long bi) {
  // total order: lower cost first, ties -> lower global index (the serial scan of a CPU sweep)
  return (c < bc) || (c == bc && i < bi);
}

__device__ __forceinline__ void warp_reduce(double& c, long long& i) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    const double oc = __shfl_xor_sync(0xffffffffu, c, m);
    const long long oi = __shfl_xor_sync(0xffffffffu, i, m);
    if (better(oc, oi, c, i)) {
      c = oc;
      i = oi;
    }
  }
}

}  // namespace mtg
#endif

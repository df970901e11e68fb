// libmtg_cuda.so — solve_canonical_kernel instantiations, AOS layout.
#include "solve_launch.cuh"

MTG_REGISTER_TABLES()

namespace mtg {
int launch_solve_canonical_aos(mtg_ctx* ctx, int N, int D, const SolveCanonicalParams& p, cudaStream_t s) {
  return solve_launch::launch_solve_canonical_n<true>(ctx, N, D, p, s);
}
}  // namespace mtg

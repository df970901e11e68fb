// Process-wide handle of libmtg_cuda.so for the class API, and the CHECK macro that reproduces the
// reference's error convention for programmer errors: glog CHECK => message + abort
// (SURVEY.md section 8b). There is no CPU path: without a CUDA device the first compute call aborts.
#ifndef MTG_SHIM_RUNTIME_H_
#define MTG_SHIM_RUNTIME_H_

#include <cstdio>
#include <cstdlib>

#include "../mtg_cuda.h"

#define MTG_SHIM_CHECK(cond, what)                                                             \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      std::fprintf(stderr, "Check failed: %s (%s) at %s:%d\n", #cond, what, __FILE__, __LINE__); \
      std::abort();                                                                            \
    }                                                                                          \
  } while (0)

namespace mav_trajectory_generation {
namespace runtime {

struct Holder {
  mtg_ctx* ctx = nullptr;
  ~Holder() {
    if (ctx) mtg_destroy(ctx);
  }
};

// one context per process (device from MTG_DEVICE, default 0); like the reference's objects the
// API is not internally synchronised
inline mtg_ctx* context() {
  static Holder h;
  if (!h.ctx) {
    const char* dev = std::getenv("MTG_DEVICE");
    const int rc = mtg_create(dev ? std::atoi(dev) : 0, &h.ctx);
    if (rc != MTG_OK) {
      std::fprintf(stderr,
                   "mav_tube_trajectory_generation (B200): mtg_create failed (rc = %d): a CUDA device is "
                   "required, there is no CPU fallback\n",
                   rc);
      std::abort();
    }
  }
  return h.ctx;
}

inline void check_rc(int rc, const char* what) {
  if (rc != MTG_OK) {
    std::fprintf(stderr, "%s failed (rc = %d): %s\n", what, rc, mtg_last_error(context()));
    std::abort();
  }
}

inline mtg_problem_desc desc(int B, int K, int D, int N, int derivative_to_optimize) {
  mtg_problem_desc d;
  d.B = B;
  d.K = K;
  d.D = D;
  d.N = N;
  d.derivative_to_optimize = derivative_to_optimize;
  d.memory = MTG_MEM_HOST;
  d.layout = MTG_LAYOUT_AOS;  // the reference's per-object order
  return d;
}

}  // namespace runtime
}  // namespace mav_trajectory_generation
#endif

// Device-side constant tables shared by all kernels of libmtg_cuda.so.
#ifndef MTG_DEVICE_TABLES_CUH_
#define MTG_DEVICE_TABLES_CUH_

#include "tables.h"

namespace mtg {

// Tables of one (N, derivative_to_optimize): only the solve / cost kernels read them.
struct DevTables {
  double H1[MTG_TAB_LD * MTG_TAB_LD];
  double Ainv1[MTG_TAB_LD * MTG_TAB_LD];
  double W[MTG_TAB_LD * MTG_TAB_LD];      // H1 = W^T W, (N-d) x N
  double Lt[MTG_TAB_LD * MTG_TAB_LD];     // W = Lt Ainv1[d.., :], upper triangular (N-d) x (N-d)
  double inv_factorial[MTG_TAB_LD];  // 1/B(j,j) = 1/j!  (the A(0) diagonal inverse, LIN_I:152-155)
  int N;
  int derivative;
};

// Polynomial::base_coefficients_ (polynomial.cpp:145-161): independent of (N, derivative), uploaded
// ONCE per device when the first context on it is created and never written again, so the
// evaluation / extrema kernels share no mutable state with anything.
struct DevBase {
  double base[MTG_BASE_LD * MTG_BASE_LD];
};

#ifdef __CUDACC__
// Every kernel translation unit of libmtg_cuda.so has its own private copies (no
// relocatable device code); host_common.h registers an uploader per unit.
// TableGuard (core.cu) switches c_tab; c_base is write-once.
static __constant__ DevTables c_tab;
static __constant__ DevBase c_base;
#endif

}  // namespace mtg
#endif

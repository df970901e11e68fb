// Device-side constant tables shared by all kernels of libmtg_cuda.so.
#ifndef MTG_DEVICE_TABLES_CUH_
#define MTG_DEVICE_TABLES_CUH_

#include "tables.h"

namespace mtg {

struct DevTables {
  double H1[MTG_TAB_LD * MTG_TAB_LD];
  double Ainv1[MTG_TAB_LD * MTG_TAB_LD];
  double W[MTG_TAB_LD * MTG_TAB_LD];      // H1 = W^T W, (N-d) x N
  double Lt[MTG_TAB_LD * MTG_TAB_LD];     // W = Lt Ainv1[d.., :], upper triangular (N-d) x (N-d)
  double base[MTG_BASE_LD * MTG_BASE_LD];
  double inv_factorial[MTG_TAB_LD];  // 1/B(j,j) = 1/j!  (the A(0) diagonal inverse, LIN_I:152-155)
  int N;
  int derivative;
};

#ifdef __CUDACC__
// Every kernel translation unit of libmtg_cuda.so has its own private copy (no
// relocatable device code); host_common.h registers an uploader per unit and
// ensure_tables() refreshes all copies together.
static __constant__ DevTables c_tab;
#endif

}  // namespace mtg
#endif

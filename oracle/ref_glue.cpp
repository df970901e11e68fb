// TEST INFRASTRUCTURE (oracle/): C entry points onto the UNMODIFIED reference
// Jenkins-Traub translation unit, compiled from /root/reference by
// oracle/Makefile into oracle/_ref/librpoly_ref.so. Nothing here is product
// code; only tests/, __graft_entry__.smoke() and bench.py's CPU legs load it.
#include <cstring>

#include "mav_tube_trajectory_generation/rpoly/rpoly_ak1.h"

namespace mav_trajectory_generation {
// Defined (non-static) at reference src/rpoly/rpoly_ak1.cpp:942-946.
void rpolyWrapper(double* coefficients_decreasing, int* degree,
                  double* roots_real, double* roots_imag);
}  // namespace mav_trajectory_generation

extern "C" {

// Raw core: coefficients in DECREASING powers, degree in/out
// (rpoly_ak1.cpp:153-389 semantics, arrays must hold 101 / 100 doubles).
void mtg_ref_rpoly(double* coefficients_decreasing, int* degree,
                   double* roots_real, double* roots_imag) {
  mav_trajectory_generation::rpolyWrapper(coefficients_decreasing, degree,
                                          roots_real, roots_imag);
}

// The reference's own wrapper (rpoly_ak1.cpp:70-117) driven through the shim
// vector types: coefficients in INCREASING powers. Returns the wrapper's bool;
// *n_roots receives roots->size().
int mtg_ref_find_roots_jenkins_traub(const double* coefficients_increasing,
                                     int n, double* roots_real,
                                     double* roots_imag, int* n_roots) {
  Eigen::VectorXd c(coefficients_increasing, n);
  Eigen::VectorXcd roots;
  const bool ok = mav_trajectory_generation::findRootsJenkinsTraub(c, &roots);
  *n_roots = roots.size();
  for (int i = 0; i < roots.size(); ++i) {
    roots_real[i] = roots[i].real();
    roots_imag[i] = roots[i].imag();
  }
  return ok ? 1 : 0;
}

int mtg_ref_find_last_non_zero_coeff(const double* coefficients, int n) {
  Eigen::VectorXd c(coefficients, n);
  return mav_trajectory_generation::findLastNonZeroCoeff(c);
}

}  // extern "C"

// TEST INFRASTRUCTURE (oracle/): include shim, NOT reference code.
//
// The reference's Jenkins-Traub translation unit (src/rpoly/rpoly_ak1.cpp)
// includes "mav_tube_trajectory_generation/rpoly/rpoly_ak1.h", which in the
// reference tree pulls <Eigen/Eigen>. Eigen is not installed on this image,
// so oracle/Makefile puts THIS directory first on the include path and
// compiles the reference .cpp *where it lies* under /root/reference. The shim
// declares just enough of the two Eigen vector types for the 60-line wrapper
// (rpoly_ak1.cpp:57-117) to compile; the numerical core (rpoly_ak1.cpp:119-940)
// needs only <cmath>/<cfloat> and is untouched.
#ifndef MTG_ORACLE_REF_SHIM_RPOLY_AK1_H_
#define MTG_ORACLE_REF_SHIM_RPOLY_AK1_H_

#include <complex>
#include <cstddef>
#include <limits>
#include <vector>

namespace Eigen {

class VectorXd {
 public:
  VectorXd() {}
  explicit VectorXd(int n) : v_(static_cast<size_t>(n), 0.0) {}
  VectorXd(const double* p, int n) : v_(p, p + n) {}
  int size() const { return static_cast<int>(v_.size()); }
  double operator()(int i) const { return v_[static_cast<size_t>(i)]; }
  double& operator()(int i) { return v_[static_cast<size_t>(i)]; }
  VectorXd head(int n) const { return VectorXd(v_.data(), n); }
  VectorXd reverse() const {
    VectorXd r(size());
    for (int i = 0; i < size(); ++i) r(i) = v_[v_.size() - 1 - static_cast<size_t>(i)];
    return r;
  }

 private:
  std::vector<double> v_;
};

class VectorXcd {
 public:
  void resize(int n) { v_.assign(static_cast<size_t>(n), std::complex<double>()); }
  int size() const { return static_cast<int>(v_.size()); }
  std::complex<double>& operator[](int i) { return v_[static_cast<size_t>(i)]; }
  const std::complex<double>& operator[](int i) const { return v_[static_cast<size_t>(i)]; }

 private:
  std::vector<std::complex<double> > v_;
};

}  // namespace Eigen

namespace mav_trajectory_generation {

int findLastNonZeroCoeff(const Eigen::VectorXd& coefficients);

bool findRootsJenkinsTraub(const Eigen::VectorXd& coefficients_increasing,
                           Eigen::VectorXcd* roots);

}  // namespace mav_trajectory_generation

#endif  // MTG_ORACLE_REF_SHIM_RPOLY_AK1_H_

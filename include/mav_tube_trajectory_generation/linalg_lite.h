// Minimal dense vector / matrix types for the class API when Eigen is not installed (it is not on
// the build image). With Eigen present (__has_include(<Eigen/Core>)) the real types are used, so the
// signatures are exactly the reference's (Eigen::VectorXd in, Eigen::VectorXd out).
#ifndef MTG_SHIM_LINALG_LITE_H_
#define MTG_SHIM_LINALG_LITE_H_

#include <cmath>
#include <cstddef>
#include <vector>

#if defined(__has_include)
#if __has_include(<Eigen/Core>) && !defined(MTG_SHIM_NO_EIGEN)
#define MTG_SHIM_HAVE_EIGEN 1
#include <Eigen/Core>
#endif
#endif

namespace mav_trajectory_generation {

#ifdef MTG_SHIM_HAVE_EIGEN
typedef Eigen::VectorXd VectorXd;
typedef Eigen::MatrixXd MatrixXd;
#else
class VectorXd {
 public:
  VectorXd() {}
  explicit VectorXd(std::size_t n) : v_(n, 0.0) {}
  VectorXd(std::initializer_list<double> init) : v_(init) {}
  static VectorXd Zero(std::size_t n) { return VectorXd(n); }
  static VectorXd Constant(std::size_t n, double value) {
    VectorXd r(n);
    for (double& x : r.v_) x = value;
    return r;
  }
  std::size_t size() const { return v_.size(); }
  void resize(std::size_t n) { v_.assign(n, 0.0); }
  void setZero() { v_.assign(v_.size(), 0.0); }
  double& operator[](std::size_t i) { return v_[i]; }
  double operator[](std::size_t i) const { return v_[i]; }
  double& operator()(std::size_t i) { return v_[i]; }
  double operator()(std::size_t i) const { return v_[i]; }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  double norm() const {
    double s = 0.0;
    for (double x : v_) s += x * x;
    return std::sqrt(s);
  }
  bool operator==(const VectorXd& o) const { return v_ == o.v_; }
  bool operator!=(const VectorXd& o) const { return !(v_ == o.v_); }

 private:
  std::vector<double> v_;
};

// row-major dense matrix
class MatrixXd {
 public:
  MatrixXd() : r_(0), c_(0) {}
  MatrixXd(std::size_t rows, std::size_t cols) : r_(rows), c_(cols), v_(rows * cols, 0.0) {}
  static MatrixXd Zero(std::size_t rows, std::size_t cols) { return MatrixXd(rows, cols); }
  void resize(std::size_t rows, std::size_t cols) {
    r_ = rows;
    c_ = cols;
    v_.assign(rows * cols, 0.0);
  }
  void setZero() { v_.assign(v_.size(), 0.0); }
  std::size_t rows() const { return r_; }
  std::size_t cols() const { return c_; }
  double& operator()(std::size_t i, std::size_t j) { return v_[i * c_ + j]; }
  double operator()(std::size_t i, std::size_t j) const { return v_[i * c_ + j]; }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }

 private:
  std::size_t r_, c_;
  std::vector<double> v_;
};
#endif

}  // namespace mav_trajectory_generation
#endif

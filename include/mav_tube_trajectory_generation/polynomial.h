// Polynomial with coefficients in INCREASING powers (polynomial.h:35-36). Containers and the
// integer derivative table live on the host; evaluate() is the E1 row of the hot path and runs
// through the C ABI (a single-segment, single-dimension trajectory, B = 1) — there is no CPU
// evaluation path in this library. Batched callers use mtg_eval_*_batch directly.
#ifndef MTG_SHIM_POLYNOMIAL_H_
#define MTG_SHIM_POLYNOMIAL_H_

#include <cmath>
#include <limits>
#include <utility>
#include <vector>

#include "linalg_lite.h"
#include "mtg_runtime.h"

namespace mav_trajectory_generation {

class Polynomial {
 public:
  typedef std::vector<Polynomial> Vector;
  static constexpr int kMaxN = 12;              // polynomial.h:45
  static constexpr int kMaxConvolutionSize = 2 * kMaxN - 2;  // polynomial.h:48

  explicit Polynomial(int N) : N_(N), coefficients_(N) {}
  Polynomial(int N, const VectorXd& coeffs) : N_(N), coefficients_(coeffs) {
    MTG_SHIM_CHECK(N_ == (int)coeffs.size(), "number of coefficients != N");  // polynomial.h:58
  }
  explicit Polynomial(const VectorXd& coeffs) : N_((int)coeffs.size()), coefficients_(coeffs) {}

  int N() const { return N_; }
  bool operator==(const Polynomial& rhs) const { return N_ == rhs.N_ && coefficients_ == rhs.coefficients_; }
  bool operator!=(const Polynomial& rhs) const { return !operator==(rhs); }

  void setCoefficients(const VectorXd& coeffs) {
    MTG_SHIM_CHECK(N_ == (int)coeffs.size(), "number of coefficients != N");
    coefficients_ = coeffs;
  }

  // B(n, i) = i! / (i - n)!  (src/polynomial.cpp:145-161)
  static double baseCoefficient(int derivative, int i) {
    if (i < derivative) return 0.0;
    double r = 1.0;
    for (int q = i - derivative + 1; q <= i; ++q) r *= q;
    return r;
  }

  // coefficients of the derivative, still N long with the tail zero (polynomial.h:99-113)
  VectorXd getCoefficients(int derivative = 0) const {
    MTG_SHIM_CHECK(derivative <= N_ && derivative >= 0, "derivative out of range");
    VectorXd r = VectorXd::Zero(N_);
    for (int j = 0; j + derivative < N_; ++j)
      r[j] = baseCoefficient(derivative, j + derivative) * coefficients_[j + derivative];
    return r;
  }

  // polynomial.h:136-149 on the GPU
  double evaluate(double t, int derivative = 0) const {
    if (derivative >= N_) return 0.0;
    const double T = std::fmax(2.0 * std::fabs(t), 1.0);   // any duration that contains t
    mtg_problem_desc d = runtime::desc(1, 1, 1, N_, 0);
    std::vector<double> c(coefficients_.data(), coefficients_.data() + N_);
    if (N_ & 1) {  // the kernels take even N: pad with a zero leading coefficient
      c.push_back(0.0);
      d.N = N_ + 1;
    }
    double out = 0.0;
    runtime::check_rc(mtg_eval_at_batch(runtime::context(), &d, c.data(), &T, &t, 1, derivative, &out, nullptr,
                                        nullptr, nullptr),
                      "mtg_eval_at_batch");
    return out;
  }
  // value and derivatives 0 .. result->size()-1 (polynomial.h:118-132)
  void evaluate(double t, VectorXd* result) const {
    MTG_SHIM_CHECK(result != nullptr, "result is null");
    for (size_t k = 0; k < (size_t)result->size(); ++k) (*result)[k] = evaluate(t, (int)k);
  }

  // row of the mapping matrix: derivative-th derivative of the monomial basis at t (polynomial.h:201-228)
  static void baseCoeffsWithTime(int N, int derivative, double t, VectorXd* coeffs) {
    MTG_SHIM_CHECK(derivative < N && derivative >= 0, "derivative out of range");
    MTG_SHIM_CHECK(coeffs != nullptr, "coeffs is null");
    *coeffs = VectorXd::Zero(N);
    (*coeffs)[derivative] = baseCoefficient(derivative, derivative);
    if (std::fabs(t) < std::numeric_limits<double>::epsilon()) return;
    double tp = t;
    for (int j = derivative + 1; j < N; ++j) {
      (*coeffs)[j] = baseCoefficient(derivative, j) * tp;
      tp *= t;
    }
  }

  // Real roots of the derivative-th derivative inside [t_start, t_end], ascending — the part of getRoots
  // (findRootsJenkinsTraub, src/polynomial.cpp:28-30) that the library consumes (only real, in-range,
  // sign-changing roots; see mtg_poly_real_roots_batch).
  bool getRealRoots(int derivative, double t_start, double t_end, std::vector<double>* roots) const {
    MTG_SHIM_CHECK(roots != nullptr, "roots is null");
    roots->clear();
    const VectorXd c = getCoefficients(derivative);
    const int n = N_ - derivative;
    if (n < 1) return true;
    std::vector<double> cc(c.data(), c.data() + n), r(n);
    int32_t nr = 0;
    uint32_t st = 0;
    runtime::check_rc(mtg_poly_real_roots_batch(runtime::context(), 1, n, MTG_MEM_HOST, MTG_LAYOUT_AOS, cc.data(),
                                                &t_start, &t_end, n, r.data(), &nr, &st, nullptr),
                      "mtg_poly_real_roots_batch");
    roots->assign(r.begin(), r.begin() + nr);
    return true;
  }
  // src/polynomial.cpp:65-83: [t_start, t_end, real roots of p^(derivative + 1) in range]
  bool computeMinMaxCandidates(double t_start, double t_end, int derivative, std::vector<double>* candidates) const {
    MTG_SHIM_CHECK(candidates != nullptr, "candidates is null");
    candidates->clear();
    if (N_ - derivative - 1 < 0) return false;
    if (t_start > t_end) return false;  // selectMinMaxCandidatesFromRoots, :37-40
    std::vector<double> roots;
    getRealRoots(derivative + 1, t_start, t_end, &roots);
    candidates->push_back(t_start);
    candidates->push_back(t_end);
    candidates->insert(candidates->end(), roots.begin(), roots.end());
    return true;
  }
  // src/polynomial.cpp:116-143: first candidate wins ties
  bool selectMinMaxFromCandidates(const std::vector<double>& candidates, int derivative,
                                  std::pair<double, double>* minimum, std::pair<double, double>* maximum) const {
    MTG_SHIM_CHECK(minimum != nullptr && maximum != nullptr, "null output");
    if (candidates.empty()) return false;
    minimum->first = maximum->first = candidates[0];
    minimum->second = std::numeric_limits<double>::max();
    maximum->second = std::numeric_limits<double>::lowest();
    for (double t : candidates) {
      const double v = evaluate(t, derivative);
      if (v < minimum->second) *minimum = std::make_pair(t, v);
      if (v > maximum->second) *maximum = std::make_pair(t, v);
    }
    return true;
  }
  // src/polynomial.cpp:99-114
  bool computeMinMax(double t_start, double t_end, int derivative, std::pair<double, double>* minimum,
                     std::pair<double, double>* maximum) const {
    std::vector<double> candidates;
    if (!computeMinMaxCandidates(t_start, t_end, derivative, &candidates)) return false;
    return selectMinMaxFromCandidates(candidates, derivative, minimum, maximum);
  }

  static int getConvolutionLength(int data_size, int kernel_size) { return data_size + kernel_size - 1; }
  // src/polynomial.cpp:163-181 (host helper; the extrema kernel convolves on the device)
  static VectorXd convolve(const VectorXd& data, const VectorXd& kernel) {
    const int n = getConvolutionLength((int)data.size(), (int)kernel.size());
    VectorXd out = VectorXd::Zero(n);
    for (int i = 0; i < (int)data.size(); ++i)
      for (int j = 0; j < (int)kernel.size(); ++j) out[i + j] += data[i] * kernel[j];
    return out;
  }

 private:
  int N_;
  VectorXd coefficients_;
};

}  // namespace mav_trajectory_generation
#endif

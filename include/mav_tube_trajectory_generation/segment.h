// Segment: D polynomials + a duration (segment.h:43-125). evaluate() and the magnitude extrema go
// through the C ABI as a one-segment trajectory (B = 1).
#ifndef MTG_SHIM_SEGMENT_H_
#define MTG_SHIM_SEGMENT_H_

#include <cstdint>
#include <vector>

#include "extremum.h"
#include "motion_defines.h"
#include "polynomial.h"

namespace mav_trajectory_generation {

class Segment {
 public:
  typedef std::vector<Segment> Vector;

  Segment(int N, int D) : time_(0.0), N_(N), D_(D) { polynomials_.resize(D_, Polynomial(N_)); }

  bool operator==(const Segment& rhs) const {
    return N_ == rhs.N_ && D_ == rhs.D_ && time_ == rhs.time_ && polynomials_ == rhs.polynomials_;
  }
  bool operator!=(const Segment& rhs) const { return !operator==(rhs); }

  int D() const { return D_; }
  int N() const { return N_; }
  double getTime() const { return time_; }
  uint64_t getTimeNSec() const { return static_cast<uint64_t>(1.0e9 * time_); }
  void setTime(double time_sec) { time_ = time_sec; }
  void setTimeNSec(uint64_t time_ns) { time_ = time_ns * 1.0e-9; }

  Polynomial& operator[](size_t idx) {
    MTG_SHIM_CHECK(idx < (size_t)D_, "dimension index out of range");  // SEG_C:42
    return polynomials_[idx];
  }
  const Polynomial& operator[](size_t idx) const {
    MTG_SHIM_CHECK(idx < (size_t)D_, "dimension index out of range");  // SEG_C:47
    return polynomials_[idx];
  }
  const Polynomial::Vector& getPolynomialsRef() const { return polynomials_; }

  // coefficients [D][N] (even N: zero padded) for the C ABI
  void packCoefficients(int n_padded, std::vector<double>* out) const {
    for (int d = 0; d < D_; ++d) {
      const VectorXd c = polynomials_[d].getCoefficients(0);
      for (int j = 0; j < n_padded; ++j) out->push_back(j < N_ ? c[j] : 0.0);
    }
  }

  // segment.cpp:51-58
  VectorXd evaluate(double t, int derivative = derivative_order::POSITION) const {
    const int n = N_ + (N_ & 1);
    std::vector<double> c;
    packCoefficients(n, &c);
    const double T = std::fmax(std::fmax(2.0 * std::fabs(t), time_), 1.0);
    mtg_problem_desc d = runtime::desc(1, 1, D_, n, 0);
    VectorXd out(D_);
    runtime::check_rc(mtg_eval_at_batch(runtime::context(), &d, c.data(), &T, &t, 1, derivative, out.data(),
                                        nullptr, nullptr, nullptr),
                      "mtg_eval_at_batch");
    return out;
  }

  // segment.cpp:82-133 / 135-158: candidate times [t_start, t_end, roots of the magnitude derivative in range]
  // and their magnitudes over `dimensions`, through mtg_extrema_candidates_batch (one-segment trajectory)
  bool computeMinMaxMagnitudeCandidates(int derivative, double t_start, double t_end,
                                        const std::vector<int>& dimensions, std::vector<Extremum>* candidates) const {
    MTG_SHIM_CHECK(candidates != nullptr, "candidates is null");
    candidates->clear();
    if (dimensions.empty()) return false;  // segment.cpp:89-91
    int mask = 0;
    for (int dim : dimensions) {
      if (dim < 0 || dim >= D_) return false;  // segment.cpp:97-102
      mask |= 1 << dim;
    }
    const int n = N_ + (N_ & 1);
    std::vector<double> c;
    packCoefficients(n, &c);
    const int MC = 2 * n;
    std::vector<double> ct(MC), cv(MC);
    int32_t nc = 0;
    mtg_problem_desc d = runtime::desc(1, 1, D_, n, 0);
    runtime::check_rc(mtg_extrema_candidates_batch(runtime::context(), &d, c.data(), &time_, &t_start, &t_end,
                                                   derivative, mask, MC, ct.data(), cv.data(), &nc, nullptr, nullptr),
                      "mtg_extrema_candidates_batch");
    for (int q = 0; q < nc && q < MC; ++q) candidates->push_back(Extremum(ct[q], cv[q], 0));
    return true;
  }
  bool computeMinMaxMagnitudeCandidateTimes(int derivative, double t_start, double t_end,
                                            const std::vector<int>& dimensions,
                                            std::vector<double>* candidate_times) const {
    MTG_SHIM_CHECK(candidate_times != nullptr, "candidate_times is null");
    candidate_times->clear();
    std::vector<Extremum> cand;
    if (!computeMinMaxMagnitudeCandidates(derivative, t_start, t_end, dimensions, &cand)) return false;
    for (const Extremum& e : cand) candidate_times->push_back(e.time);
    return true;
  }

  // segment.cpp:160-184 (pure selection, host)
  bool selectMinMaxMagnitudeFromCandidates(int /*derivative*/, double t_start, double t_end,
                                           const std::vector<int>& /*dimensions*/,
                                           const std::vector<Extremum>& candidates, Extremum* minimum,
                                           Extremum* maximum) const {
    MTG_SHIM_CHECK(minimum != nullptr && maximum != nullptr, "null output");
    if (t_start > t_end) return false;
    minimum->value = std::numeric_limits<double>::max();
    maximum->value = std::numeric_limits<double>::lowest();
    for (const Extremum& c : candidates) {
      if (c.time < t_start || c.time > t_end) continue;
      if (*maximum < c) *maximum = c;
      if (c < *minimum) *minimum = c;
    }
    return true;
  }

  bool getSegmentWithSingleDimension(int dimension, Segment* new_segment) const {
    if (dimension < 0 || dimension >= D_) return false;
    *new_segment = Segment(N_, 1);
    (*new_segment)[0] = polynomials_[dimension];
    new_segment->setTime(time_);
    return true;
  }
  bool getSegmentWithAppendedDimension(const Segment& other, Segment* new_segment) const {
    if (N_ == 0 || D_ == 0) {
      *new_segment = other;
      return true;
    }
    if (other.N_ == 0 || other.D_ == 0) {
      *new_segment = *this;
      return true;
    }
    if (time_ != other.time_ || N_ != other.N_) return false;  // (the reference also pads unequal N; not needed here)
    *new_segment = Segment(N_, D_ + other.D_);
    for (int d = 0; d < D_; ++d) (*new_segment)[d] = polynomials_[d];
    for (int d = 0; d < other.D_; ++d) (*new_segment)[D_ + d] = other.polynomials_[d];
    new_segment->setTime(time_);
    return true;
  }

 protected:
  Polynomial::Vector polynomials_;
  double time_;

 private:
  int N_;
  int D_;
};

}  // namespace mav_trajectory_generation
#endif

// Trajectory: K segments (trajectory.h:32-130). evaluate / evaluateRange / computeMinMaxMagnitude
// are rows E3 / E4 / E6 of the hot path and run through the C ABI with B = 1; batched callers hand
// whole batches to mtg_eval_range_batch / mtg_feasibility_batch / mtg_extrema_batch directly.
#ifndef MTG_SHIM_TRAJECTORY_H_
#define MTG_SHIM_TRAJECTORY_H_

#include <cmath>
#include <cstdio>
#include <vector>

#include "extremum.h"
#include "segment.h"
#include "vertex.h"

namespace mav_trajectory_generation {

class Trajectory {
 public:
  Trajectory() : D_(0), N_(0), max_time_(0.0) {}

  bool operator==(const Trajectory& rhs) const {
    return D_ == rhs.D_ && N_ == rhs.N_ && max_time_ == rhs.max_time_ && segments_ == rhs.segments_;
  }
  bool operator!=(const Trajectory& rhs) const { return !operator==(rhs); }

  int D() const { return D_; }
  int N() const { return N_; }
  int K() const { return (int)segments_.size(); }
  bool empty() const { return segments_.empty(); }
  void clear() {
    segments_.clear();
    D_ = N_ = 0;
    max_time_ = 0.0;
  }

  void setSegments(const Segment::Vector& segments) {
    MTG_SHIM_CHECK(!segments.empty(), "segments must not be empty");  // TRAJ_H:55
    D_ = segments.front().D();
    N_ = segments.front().N();
    max_time_ = 0.0;
    segments_.clear();
    addSegments(segments);
  }
  void addSegments(const Segment::Vector& segments) {
    for (const Segment& s : segments) {
      MTG_SHIM_CHECK(s.D() == D_, "segment dimension");  // TRAJ_H:67-68
      MTG_SHIM_CHECK(s.N() == N_, "segment coefficient count");
      max_time_ += s.getTime();  // summed in segment order: what mtg_max_time_batch reproduces
    }
    segments_.insert(segments_.end(), segments.begin(), segments.end());
  }
  void getSegments(Segment::Vector* segments) const {
    MTG_SHIM_CHECK(segments != nullptr, "segments is null");
    *segments = segments_;
  }
  const Segment::Vector& segments() const { return segments_; }

  double getMinTime() const { return 0.0; }
  double getMaxTime() const { return max_time_; }
  std::vector<double> getSegmentTimes() const {
    std::vector<double> t(segments_.size());
    for (size_t i = 0; i < t.size(); ++i) t[i] = segments_[i].getTime();
    return t;
  }

  Trajectory getTrajectoryWithSingleDimension(int dimension) const {
    MTG_SHIM_CHECK(dimension < D_, "dimension out of range");
    Segment::Vector segs;
    for (const Segment& s : segments_) {
      Segment one(N_, 1);
      one[0] = s[dimension];
      segs.push_back(one);  // (like the reference, the single-dimension copy does not carry the time)
    }
    Trajectory t;
    t.setSegments(segs);
    return t;
  }
  bool getTrajectoryWithAppendedDimension(const Trajectory& other, Trajectory* out) const {
    if (N_ == 0 || D_ == 0) {
      *out = other;
      return true;
    }
    if (other.N() == 0 || other.D() == 0) {
      *out = *this;
      return true;
    }
    MTG_SHIM_CHECK(K() == other.K(), "segment counts differ");
    Segment::Vector segs;
    for (size_t k = 0; k < segments_.size(); ++k) {
      Segment s(0, 0);
      if (!segments_[k].getSegmentWithAppendedDimension(other.segments()[k], &s)) return false;
      segs.push_back(s);
    }
    out->setSegments(segs);
    return true;
  }
  bool addTrajectories(const std::vector<Trajectory>& trajectories, Trajectory* merged) const {
    MTG_SHIM_CHECK(merged != nullptr, "merged is null");
    *merged = *this;
    for (const Trajectory& t : trajectories) {
      if (t.D() != D_ || t.N() != N_) return false;
      merged->addSegments(t.segments());
    }
    return true;
  }

  Vertex getVertexAtTime(double t, int max_derivative_order) const {
    Vertex v(D_);
    for (int i = 0; i <= max_derivative_order; ++i) v.addConstraint(i, evaluate(t, i));
    return v;
  }
  Vertex getStartVertex(int max_derivative_order) const { return getVertexAtTime(0.0, max_derivative_order); }
  Vertex getGoalVertex(int max_derivative_order) const { return getVertexAtTime(max_time_, max_derivative_order); }

  // ---- the C-ABI view of this object: coefficients [K][D][n] (n even), seg_times [K]
  int paddedN() const { return N_ + (N_ & 1); }
  void pack(std::vector<double>* coeffs, std::vector<double>* times) const {
    for (const Segment& s : segments_) {
      s.packCoefficients(paddedN(), coeffs);
      times->push_back(s.getTime());
    }
  }

  // Trajectory::evaluate, trajectory.cpp:41-72 (vertex times belong to the segment on their right;
  // t == max time evaluates the last segment; out of range: error message + zero vector)
  VectorXd evaluate(double t, int derivative = derivative_order::POSITION) const {
    MTG_SHIM_CHECK(!segments_.empty(), "empty trajectory");
    std::vector<double> c, times;
    pack(&c, &times);
    mtg_problem_desc d = runtime::desc(1, K(), D_, paddedN(), 0);
    VectorXd out(D_);
    uint32_t status = 0;
    runtime::check_rc(mtg_eval_at_batch(runtime::context(), &d, c.data(), times.data(), &t, 1, derivative,
                                        out.data(), nullptr, &status, nullptr),
                      "mtg_eval_at_batch");
    if (status & MTG_ST_OUT_OF_RANGE) std::fprintf(stderr, "Time out of range of the trajectory!\n");
    return out;
  }

  // Trajectory::evaluateRange, trajectory.cpp:74-134: the reference's serial sampling recurrence,
  // replayed bit-exactly on the device (sample count, sampling times and segment of every sample)
  void evaluateRange(double t_start, double t_end, double dt, int derivative, std::vector<VectorXd>* result,
                     std::vector<double>* sampling_times = nullptr) const {
    MTG_SHIM_CHECK(result != nullptr, "result is null");
    MTG_SHIM_CHECK(!segments_.empty(), "empty trajectory");
    result->clear();
    if (sampling_times) sampling_times->clear();
    std::vector<double> c, times;
    pack(&c, &times);
    mtg_problem_desc d = runtime::desc(1, K(), D_, paddedN(), 0);
    double hint = (t_end - t_start) / dt + 1.0;  // the reference's reserve() (trajectory.cpp:78-86)
    if (!(hint >= 0.0) || hint > 1.0e9) hint = 0.0;
    int cap = (int)hint + K() + 8;
    for (;;) {
      std::vector<double> samples((size_t)cap * D_), st((size_t)cap);
      int32_t n = 0;
      uint32_t status = 0;
      runtime::check_rc(mtg_eval_range_batch(runtime::context(), &d, c.data(), times.data(), &t_start, &t_end, &dt,
                                             derivative, cap, samples.data(), st.data(), nullptr, &n, &status,
                                             nullptr),
                        "mtg_eval_range_batch");
      if (status & MTG_ST_TRUNCATED) {
        cap *= 2;
        continue;
      }
      if (status & MTG_ST_OUT_OF_RANGE) {
        std::fprintf(stderr, "Start time out of range of the trajectory!\n");  // trajectory.cpp:104-107
        return;
      }
      for (int k = 0; k < n; ++k) {
        VectorXd v(D_);
        for (int dim = 0; dim < D_; ++dim) v[dim] = samples[(size_t)k * D_ + dim];
        result->push_back(v);
        if (sampling_times) sampling_times->push_back(st[k]);
      }
      return;
    }
  }

  // Trajectory::computeMinMaxMagnitude, trajectory.cpp:184-220
  bool computeMinMaxMagnitude(int derivative, const std::vector<int>& dimensions, Extremum* minimum,
                              Extremum* maximum) const {
    MTG_SHIM_CHECK(minimum != nullptr && maximum != nullptr, "null output");
    if (dimensions.empty() || segments_.empty()) return false;  // segment.cpp:89-91
    for (int dim : dimensions)
      if (dim < 0 || dim >= D_) return false;  // segment.cpp:100-105
    const int n = paddedN(), Dsel = (int)dimensions.size();
    MTG_SHIM_CHECK(Dsel <= 4, "at most 4 dimensions");
    std::vector<double> c, times;
    for (const Segment& s : segments_) {
      for (int dim : dimensions) {
        const VectorXd pc = s[dim].getCoefficients(0);
        for (int j = 0; j < n; ++j) c.push_back(j < N_ ? pc[j] : 0.0);
      }
      times.push_back(s.getTime());
    }
    mtg_problem_desc d = runtime::desc(1, K(), Dsel, n, 0);
    int32_t mn_s = 0, mx_s = 0;
    runtime::check_rc(mtg_extrema_batch(runtime::context(), &d, c.data(), times.data(), derivative, &minimum->value,
                                        &minimum->time, &mn_s, &maximum->value, &maximum->time, &mx_s, nullptr,
                                        nullptr, nullptr, nullptr),
                      "mtg_extrema_batch");
    minimum->segment_idx = mn_s;
    maximum->segment_idx = mx_s;
    return true;
  }

 private:
  int D_;
  int N_;
  double max_time_;
  Segment::Vector segments_;
};

}  // namespace mav_trajectory_generation
#endif

// Vertex (support point with derivative constraints), segment-time heuristics and the synthetic
// vertex generators of the reference (vertex.h:42-174, src/vertex.cpp). Host-side containers: the
// input side of the hot path. Same names, arguments and CHECK behaviour as the reference.
#ifndef MTG_SHIM_VERTEX_H_
#define MTG_SHIM_VERTEX_H_

#include <cmath>
#include <map>
#include <random>
#include <utility>
#include <vector>

#include "linalg_lite.h"
#include "motion_defines.h"
#include "mtg_runtime.h"

namespace mav_trajectory_generation {

class Vertex {
 public:
  typedef std::vector<Vertex> Vector;
  typedef VectorXd ConstraintValue;
  typedef std::pair<int, ConstraintValue> Constraint;
  typedef std::map<int, ConstraintValue> Constraints;

  explicit Vertex(size_t dimension) : D_((int)dimension) {}
  int D() const { return D_; }

  // same value in every dimension
  void addConstraint(int derivative_order, double value) {
    constraints_[derivative_order] = ConstraintValue::Constant(D_, value);
  }
  void addConstraint(int derivative_order, const VectorXd& constraint) {
    MTG_SHIM_CHECK((int)constraint.size() == D_, "constraint dimension != vertex dimension");  // VTX_C:132
    constraints_[derivative_order] = constraint;
  }
  bool removeConstraint(int derivative_order) { return constraints_.erase(derivative_order) > 0; }

  // position + zero derivatives 1..up_to_derivative (src/vertex.cpp:147-153)
  void makeStartOrEnd(const VectorXd& constraint, int up_to_derivative) {
    addConstraint(derivative_order::POSITION, constraint);
    for (int i = 1; i <= up_to_derivative; ++i) constraints_[i] = ConstraintValue::Zero(D_);
  }
  void makeStartOrEnd(double value, int up_to_derivative) {
    makeStartOrEnd(VectorXd::Constant(D_, value), up_to_derivative);
  }

  bool hasConstraint(int derivative_order) const { return constraints_.count(derivative_order) > 0; }
  bool getConstraint(int derivative_order, VectorXd* constraint) const {
    MTG_SHIM_CHECK(constraint != nullptr, "constraint is null");
    Constraints::const_iterator it = constraints_.find(derivative_order);
    if (it == constraints_.end()) return false;
    *constraint = it->second;
    return true;
  }
  Constraints::const_iterator cBegin() const { return constraints_.begin(); }
  Constraints::const_iterator cEnd() const { return constraints_.end(); }
  size_t getNumberOfConstraints() const { return constraints_.size(); }

  bool isEqualTol(const Vertex& rhs, double tol) const {
    if (constraints_.size() != rhs.constraints_.size()) return false;
    Constraints::const_iterator a = constraints_.begin(), b = rhs.constraints_.begin();
    for (; a != constraints_.end(); ++a, ++b) {
      if (a->first != b->first || a->second.size() != b->second.size()) return false;
      for (size_t i = 0; i < (size_t)a->second.size(); ++i)
        if (std::fabs(a->second[i] - b->second[i]) > tol) return false;
    }
    return true;
  }

 private:
  int D_;
  Constraints constraints_;
};

namespace detail {
inline double distance(const VectorXd& a, const VectorXd& b) {
  double s = 0.0;
  for (size_t i = 0; i < (size_t)a.size(); ++i) s += (a[i] - b[i]) * (a[i] - b[i]);
  return std::sqrt(s);
}
}  // namespace detail

// src/vertex.cpp:271-287
inline double computeTimeVelocityRamp(const VectorXd& start, const VectorXd& goal, double v_max, double a_max) {
  const double dist = detail::distance(start, goal);
  const double t_acc = v_max / a_max;             // time to reach v_max
  const double d_acc = 0.5 * v_max * t_acc;       // distance covered meanwhile
  if (dist < 2.0 * d_acc) return 2.0 * std::sqrt(dist / a_max);
  return 2.0 * t_acc + (dist - 2.0 * d_acc) / v_max;
}

// src/vertex.cpp:233-250
inline std::vector<double> estimateSegmentTimesVelocityRamp(const Vertex::Vector& vertices, double v_max,
                                                            double a_max, double time_factor = 1.0) {
  (void)time_factor;  // unused by the reference as well
  MTG_SHIM_CHECK(vertices.size() >= 2, "at least two vertices");
  std::vector<double> times;
  for (size_t i = 0; i + 1 < vertices.size(); ++i) {
    VectorXd a, b;
    vertices[i].getConstraint(derivative_order::POSITION, &a);
    vertices[i + 1].getConstraint(derivative_order::POSITION, &b);
    times.push_back(computeTimeVelocityRamp(a, b, v_max, a_max));
  }
  return times;
}

// src/vertex.cpp:252-269: t = 2 d / v_max (1 + c v_max / a_max exp(-2 d / v_max))
inline std::vector<double> estimateSegmentTimesNfabian(const Vertex::Vector& vertices, double v_max, double a_max,
                                                       double magic_fabian_constant = 6.5) {
  MTG_SHIM_CHECK(vertices.size() >= 2, "at least two vertices");
  std::vector<double> times;
  for (size_t i = 0; i + 1 < vertices.size(); ++i) {
    VectorXd a, b;
    vertices[i].getConstraint(derivative_order::POSITION, &a);
    vertices[i + 1].getConstraint(derivative_order::POSITION, &b);
    const double d = detail::distance(a, b);
    times.push_back(d / v_max * 2 * (1.0 + magic_fabian_constant * v_max / a_max * std::exp(-d / v_max * 2)));
  }
  return times;
}

inline std::vector<double> estimateSegmentTimes(const Vertex::Vector& vertices, double v_max, double a_max) {
  return estimateSegmentTimesNfabian(vertices, v_max, a_max);
}

// src/vertex.cpp:27-82: std::mt19937(seed), one uniform_real_distribution per dimension, vertices
// closer than 0.2 to their predecessor are redrawn; first / last vertex are start / end vertices.
inline Vertex::Vector createRandomVertices(int maximum_derivative, size_t n_segments, const VectorXd& pos_min,
                                           const VectorXd& pos_max, size_t seed = 0) {
  MTG_SHIM_CHECK((int)n_segments >= 1, "n_segments >= 1");
  MTG_SHIM_CHECK(pos_min.size() == pos_max.size(), "pos_min / pos_max sizes");
  MTG_SHIM_CHECK(detail::distance(pos_max, pos_min) >= 0.2, "box diagonal >= 0.2");
  MTG_SHIM_CHECK(maximum_derivative > 0, "maximum_derivative > 0");
  const size_t dim = pos_min.size();
  std::mt19937 gen(seed);
  std::vector<std::uniform_real_distribution<double> > dist(dim);
  for (size_t i = 0; i < dim; ++i) dist[i] = std::uniform_real_distribution<double>(pos_min[i], pos_max[i]);
  VectorXd last(dim);
  for (size_t i = 0; i < dim; ++i) last[i] = dist[i](gen);
  Vertex::Vector vertices;
  vertices.push_back(Vertex(dim));
  vertices.front().makeStartOrEnd(last, maximum_derivative);
  for (size_t v = 1; v <= n_segments; ++v) {
    VectorXd pos(dim);
    do {
      for (size_t d = 0; d < dim; ++d) pos[d] = dist[d](gen);
    } while (!(detail::distance(pos, last) > 0.2));
    Vertex vx(dim);
    vx.addConstraint(derivative_order::POSITION, pos);
    vertices.push_back(vx);
    last = pos;
  }
  vertices.back().makeStartOrEnd(last, maximum_derivative);
  return vertices;
}

inline Vertex::Vector createRandomVertices1D(int maximum_derivative, size_t n_segments, double pos_min,
                                             double pos_max, size_t seed = 0) {
  return createRandomVertices(maximum_derivative, n_segments, VectorXd::Constant(1, pos_min),
                              VectorXd::Constant(1, pos_max), seed);
}

// src/vertex.cpp:84-120: `rounds` laps around a square in the z = center[2] plane
inline Vertex::Vector createSquareVertices(int maximum_derivative, const VectorXd& center, double side_length,
                                           int rounds) {
  const double h = side_length / 2.0;
  const double cx[4] = {center[0] - h, center[0] - h, center[0] + h, center[0] + h};
  const double cy[4] = {center[1] - h, center[1] + h, center[1] + h, center[1] - h};
  Vertex::Vector vertices;
  for (int k = 0; k <= 4 * rounds; ++k) {
    VectorXd pos(3);
    pos[0] = cx[k % 4];
    pos[1] = cy[k % 4];
    pos[2] = center[2];
    Vertex v(3);
    if (k == 0 || k == 4 * rounds)
      v.makeStartOrEnd(pos, maximum_derivative);
    else
      v.addConstraint(derivative_order::POSITION, pos);
    vertices.push_back(v);
  }
  return vertices;
}

}  // namespace mav_trajectory_generation
#endif

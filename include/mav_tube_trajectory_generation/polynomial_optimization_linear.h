// PolynomialOptimization<N>: the unconstrained (linear) minimum-derivative optimiser of the
// reference (polynomial_optimization_linear.h:45-285, impl/..._linear_impl.h) on top of
// libmtg_cuda.so. Same class, method names, arguments and error behaviour; solveLinear(),
// setFreeConstraints(), computeCost() and computeMaximumOfMagnitude() run on the GPU through the
// C ABI (B = 1); callers with many problems use the `Batch` statics or the C ABI directly.
//
// Constraint patterns: the pattern createRandomVertices / makeStartOrEnd produce (first and last
// vertex constrain derivatives 0..N/2-1, interior vertices position only) takes the two-lane
// kernel behind mtg_solve_batch; every other pattern (any subset of derivatives 0..N/2-1 at any
// vertex, LIN_I:171-252) takes mtg_solve_generic_batch.
#ifndef MTG_SHIM_POLYNOMIAL_OPTIMIZATION_LINEAR_H_
#define MTG_SHIM_POLYNOMIAL_OPTIMIZATION_LINEAR_H_

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "extremum.h"
#include "linalg_lite.h"
#include "motion_defines.h"
#include "polynomial.h"
#include "segment.h"
#include "trajectory.h"
#include "vertex.h"

namespace mav_trajectory_generation {

template <int _N = 10>
class PolynomialOptimization {
  static_assert(_N % 2 == 0, "The number of coefficients has to be even.");
  static_assert(_N >= 4 && _N <= MTG_MAX_N, "N in {4, 6, 8, 10, 12}");

 public:
  enum { N = _N };
  static constexpr int kHighestDerivativeToOptimize = N / 2 - 1;
  typedef MatrixXd SquareMatrix;  // N x N
  typedef std::vector<SquareMatrix> SquareMatrixVector;

  explicit PolynomialOptimization(size_t dimension)
      : dimension_(dimension), derivative_to_optimize_(derivative_order::INVALID), n_vertices_(0), n_segments_(0),
        n_all_constraints_(0), n_fixed_constraints_(0), n_free_constraints_(0), cost_(0.0) {
    MTG_SHIM_CHECK(dimension >= 1 && dimension <= 4, "dimension must be 1..4");
    free_constraints_compact_.resize(dimension_);
  }
  virtual ~PolynomialOptimization() {}

  // LIN_I:46-99
  virtual bool setupFromVertices(const Vertex::Vector& vertices, const std::vector<double>& segment_times,
                                 int derivative_to_optimize = kHighestDerivativeToOptimize) {
    MTG_SHIM_CHECK(derivative_to_optimize >= 0 && derivative_to_optimize <= kHighestDerivativeToOptimize,
                   "derivative_to_optimize out of range");  // LIN_I:50-55
    MTG_SHIM_CHECK(vertices.size() >= 2, "at least two vertices");
    MTG_SHIM_CHECK(segment_times.size() + 1 == vertices.size(), "segment_times.size() + 1 == vertices.size()");  // :66-67
    const int h = N / 2;
    const size_t K = segment_times.size();
    // keep constraints of order <= N/2-1 only (LIN_I:74-95); record the fixed / free pattern
    Vertex::Vector kept;
    mask_.assign((K + 1) * h, 0);
    canonical_ = true;
    size_t n_fixed = 0;
    for (size_t v = 0; v < vertices.size(); ++v) {
      MTG_SHIM_CHECK((size_t)vertices[v].D() == dimension_, "vertex dimension");
      Vertex vx(dimension_);
      for (Vertex::Constraints::const_iterator it = vertices[v].cBegin(); it != vertices[v].cEnd(); ++it)
        if (it->first >= 0 && it->first <= kHighestDerivativeToOptimize) {
          vx.addConstraint(it->first, it->second);
          mask_[v * h + it->first] = 1;
          ++n_fixed;
        }
      const bool end = (v == 0 || v == K);
      if (!(vx.hasConstraint(derivative_order::POSITION) && vx.getNumberOfConstraints() == (size_t)(end ? h : 1)))
        canonical_ = false;
      kept.push_back(vx);
    }
    vertices_ = kept;
    derivative_to_optimize_ = derivative_to_optimize;
    n_vertices_ = vertices.size();
    n_segments_ = K;
    n_all_constraints_ = K * N;
    n_fixed_constraints_ = n_fixed;
    n_free_constraints_ = (K + 1) * h - n_fixed;
    segments_.assign(n_segments_, Segment(N, (int)dimension_));
    for (size_t d = 0; d < dimension_; ++d) free_constraints_compact_[d] = VectorXd::Zero(n_free_constraints_);
    cost_ = 0.0;
    updateSegmentTimes(segment_times);
    return true;
  }

  // LIN_I:277-304. The per-segment matrices are closed forms of T on the device,
  // so there is nothing to rebuild here.
  void updateSegmentTimes(const std::vector<double>& segment_times) {
    MTG_SHIM_CHECK(segment_times.size() == n_segments_, "number of segment times");  // :281-283
    for (size_t i = 0; i < n_segments_; ++i) {
      MTG_SHIM_CHECK(segment_times[i] > 0.0, "segment time must be > 0");  // :296
      segments_[i].setTime(segment_times[i]);
    }
    segment_times_ = segment_times;
  }

  // LIN_I:337-379 (+ constructR :306-335, updateSegmentsFromCompactConstraints :254-275)
  bool solveLinear() {
    MTG_SHIM_CHECK(derivative_to_optimize_ >= 0 && derivative_to_optimize_ <= kHighestDerivativeToOptimize,
                   "setupFromVertices first");  // :339-340
    const int K = (int)n_segments_, D = (int)dimension_;
    std::vector<double> coeffs((size_t)K * D * N), free((size_t)D * n_free_constraints_ + 1);
    uint32_t status = 0;
    mtg_problem_desc d = runtime::desc(1, K, D, N, derivative_to_optimize_);
    if (canonical_) {
      std::vector<double> pos, endd;
      packVertices(&pos, &endd);
      runtime::check_rc(mtg_solve_batch(runtime::context(), &d, pos.data(), endd.data(), segment_times_.data(),
                                        coeffs.data(), &cost_, free.data(), &status, nullptr),
                        "mtg_solve_batch");
    } else {
      std::vector<double> values;
      packValues(nullptr, &values);
      runtime::check_rc(mtg_solve_generic_batch(runtime::context(), &d, mask_.data(), values.data(),
                                                segment_times_.data(), coeffs.data(), &cost_, free.data(), &status,
                                                nullptr),
                        "mtg_solve_generic_batch");
    }
    if (status & MTG_ST_NOT_SPD)
      std::fprintf(stderr, "PolynomialOptimization (B200): R_pp is numerically not positive definite\n");
    for (int dim = 0; dim < D; ++dim)
      for (size_t q = 0; q < n_free_constraints_; ++q)
        free_constraints_compact_[dim][q] = free[(size_t)dim * n_free_constraints_ + q];
    unpackSegments(coeffs);
    return true;  // like the reference, always
  }

  // LIN_I:113-130: 0.5 sum c^T Q c of the CURRENT segments with the CURRENT segment times (after
  // updateSegmentTimes without a new solve: new Q, old coefficients — like the reference)
  double computeCost() const {
    if (n_segments_ == 0) return 0.0;
    Trajectory t;
    t.setSegments(segments_);
    std::vector<double> c, times;
    t.pack(&c, &times);
    double cost = 0.0;
    mtg_problem_desc d = runtime::desc(1, (int)n_segments_, (int)dimension_, N, derivative_to_optimize_);
    runtime::check_rc(mtg_compute_cost_batch(runtime::context(), &d, c.data(), segment_times_.data(), &cost, nullptr,
                                             nullptr),
                      "mtg_compute_cost_batch");
    return cost;
  }

  void getTrajectory(Trajectory* trajectory) const {
    MTG_SHIM_CHECK(trajectory != nullptr, "trajectory is null");
    trajectory->setSegments(segments_);
  }
  void getSegments(Segment::Vector* segments) const {
    MTG_SHIM_CHECK(segments != nullptr, "segments is null");
    *segments = segments_;
  }
  void getVertices(Vertex::Vector* vertices) const {
    MTG_SHIM_CHECK(vertices != nullptr, "vertices is null");
    *vertices = vertices_;
  }
  void getSegmentTimes(std::vector<double>* segment_times) const {
    MTG_SHIM_CHECK(segment_times != nullptr, "segment_times is null");
    *segment_times = segment_times_;
  }
  // d_p per dimension, ordered by (vertex, derivative) like the reference's std::set<Constraint> (LIN_H:289-296)
  void getFreeConstraints(std::vector<VectorXd>* free_constraints) const {
    MTG_SHIM_CHECK(free_constraints != nullptr, "free_constraints is null");
    *free_constraints = free_constraints_compact_;
  }
  // d_f per dimension: (0, 0..h-1), (v, 0) v = 1..K-1, (K, 0..h-1)
  void getFixedConstraints(std::vector<VectorXd>* fixed_constraints) const {
    MTG_SHIM_CHECK(fixed_constraints != nullptr, "fixed_constraints is null");
    const int h = N / 2, K = (int)n_segments_;
    fixed_constraints->assign(dimension_, VectorXd::Zero(n_fixed_constraints_));
    for (size_t dim = 0; dim < dimension_; ++dim) {
      size_t q = 0;
      for (int v = 0; v <= K; ++v)
        for (int k = 0; k < h; ++k) {
          if (!mask_[v * h + k]) continue;
          VectorXd c;
          vertices_[v].getConstraint(k, &c);
          (*fixed_constraints)[dim][q++] = c[dim];
        }
    }
  }
  // LIN_I:489-498 + 254-275: segments (and cost) from externally chosen free derivatives
  void setFreeConstraints(const std::vector<VectorXd>& free_constraints) {
    MTG_SHIM_CHECK(free_constraints.size() == dimension_, "one vector per dimension");  // :492-494
    for (const VectorXd& v : free_constraints)
      MTG_SHIM_CHECK((size_t)v.size() == n_free_constraints_, "number of free constraints");
    free_constraints_compact_ = free_constraints;
    const int K = (int)n_segments_, D = (int)dimension_;
    std::vector<double> coeffs((size_t)K * D * N), full;
    packValues(&free_constraints, &full);  // C [d_f; d_p]: the full endpoint derivatives of every vertex
    mtg_problem_desc d = runtime::desc(1, K, D, N, derivative_to_optimize_);
    runtime::check_rc(mtg_coeffs_from_derivatives_batch(runtime::context(), &d, full.data(), segment_times_.data(),
                                                        coeffs.data(), &cost_, nullptr, nullptr),
                      "mtg_coeffs_from_derivatives_batch");
    unpackSegments(coeffs);
  }

  // LIN_I:396-417: candidate times of one segment: [t_start, t_stop, roots of the magnitude derivative in range]
  static bool computeSegmentMaximumMagnitudeCandidates(int derivative, const Segment& segment, double t_start,
                                                       double t_stop, std::vector<double>* candidates) {
    MTG_SHIM_CHECK(candidates != nullptr, "candidates is null");
    MTG_SHIM_CHECK(N - derivative - 1 > 0, "N-Derivative-1 has to be greater 0");  // :400-401
    std::vector<int> dimensions;
    for (int i = 0; i < segment.D(); ++i) dimensions.push_back(i);
    return segment.computeMinMaxMagnitudeCandidateTimes(derivative, t_start, t_stop, dimensions, candidates);
  }
  template <int Derivative>
  static bool computeSegmentMaximumMagnitudeCandidates(const Segment& segment, double t_start, double t_stop,
                                                       std::vector<double>* candidates) {
    return computeSegmentMaximumMagnitudeCandidates(Derivative, segment, t_start, t_stop, candidates);
  }
  // LIN_I:419-453: candidates by sampling (direction changes of the magnitude with |p^(d+1)| < 1e-2); the
  // samples come from one evaluateRange-free sweep: the t += dt accumulation of the reference's loop is kept
  template <int Derivative>
  static void computeSegmentMaximumMagnitudeCandidatesBySampling(const Segment& segment, double t_start,
                                                                 double t_stop, double dt,
                                                                 std::vector<double>* candidates) {
    MTG_SHIM_CHECK(candidates != nullptr, "candidates is null");
    auto norm = [](const VectorXd& v) {
      double s = 0.0;
      for (int i = 0; i < (int)v.size(); ++i) s += v[i] * v[i];
      return std::sqrt(s);
    };
    const VectorXd value_start = segment.evaluate(t_start - dt, Derivative);
    VectorXd value_old = segment.evaluate(t_start, Derivative);
    double direction = norm(value_old) - norm(value_start);
    for (double t = t_start + dt; t < t_stop + dt; t += dt) {
      const VectorXd value_new = segment.evaluate(t, Derivative);
      const double direction_new = norm(value_new) - norm(value_old);
      if (std::signbit(direction) != std::signbit(direction_new)) {
        if (norm(segment.evaluate(t - dt, Derivative + 1)) < 1e-2) candidates->push_back(t - dt);
      }
      value_old = value_new;
      direction = direction_new;
    }
  }

  // LIN_I:455-487: maximum of |p^(derivative)| over the trajectory (time relative to its segment). `candidates`,
  // if given, receives every candidate like the reference: per segment t = 0, then [0, T, roots...], and the
  // end of the last segment once more.
  Extremum computeMaximumOfMagnitude(int derivative, std::vector<Extremum>* candidates) const {
    MTG_SHIM_CHECK(N - derivative - 1 > 0, "N - derivative - 1 has to be greater 0");  // :400-401
    if (candidates) candidates->clear();
    Trajectory t;
    t.setSegments(segments_);
    std::vector<double> c, times;
    t.pack(&c, &times);
    const int K = (int)n_segments_, MC = 2 * N;
    std::vector<double> ct((size_t)K * MC), cv((size_t)K * MC);
    std::vector<int32_t> nc(K);
    mtg_problem_desc d = runtime::desc(1, K, (int)dimension_, N, derivative_to_optimize_);
    runtime::check_rc(mtg_extrema_candidates_batch(runtime::context(), &d, c.data(), times.data(), nullptr, nullptr,
                                                   derivative, 0, MC, ct.data(), cv.data(), nc.data(), nullptr, nullptr),
                      "mtg_extrema_candidates_batch");
    Extremum best;
    for (int s = 0; s < K; ++s) {
      // extrema_times = {0.0} + the segment's candidates (the call clears the vector first in the reference,
      // so the list is just [0, T, roots...]; LIN_I:466-470 with segment.cpp:86)
      for (int q = 0; q < nc[s] && q < MC; ++q) {
        const Extremum cand(ct[(size_t)s * MC + q], cv[(size_t)s * MC + q], s);
        if (best < cand) best = cand;
        if (candidates) candidates->push_back(cand);
      }
    }
    if (K > 0) {  // LIN_I:479-484: the last time of the last segment (candidate 1 of that segment)
      const Extremum cand(ct[(size_t)(K - 1) * MC + 1], cv[(size_t)(K - 1) * MC + 1], K - 1);
      if (best < cand) best = cand;
      if (candidates) candidates->push_back(cand);
    }
    return best;
  }
  template <int Derivative>
  Extremum computeMaximumOfMagnitude(std::vector<Extremum>* candidates) const {
    return computeMaximumOfMagnitude(Derivative, candidates);
  }

  // ---- small dense helpers of the reference API (host; not on the batched path)
  // LIN_I:101-111: A = [derivative bases at 0; derivative bases at T]
  static void setupMappingMatrix(double segment_time, SquareMatrix* A) {
    MTG_SHIM_CHECK(A != nullptr, "A is null");
    A->resize(N, N);
    for (int i = 0; i < N / 2; ++i) {
      VectorXd row;
      Polynomial::baseCoeffsWithTime(N, i, 0.0, &row);
      for (int j = 0; j < N; ++j) (*A)(i, j) = row[j];
      Polynomial::baseCoeffsWithTime(N, i, segment_time, &row);
      for (int j = 0; j < N; ++j) (*A)(i + N / 2, j) = row[j];
    }
  }
  // LIN_I:132-169: [diag^-1, 0; -D^-1 C diag^-1, D^-1] with the lower-right block D inverted densely
  static void invertMappingMatrix(const SquareMatrix& A, SquareMatrix* Ainv) {
    MTG_SHIM_CHECK(Ainv != nullptr, "Ainv is null");
    const int h = N / 2;
    Ainv->resize(N, N);
    std::vector<double> M((size_t)h * 2 * h, 0.0);  // [D | I] -> [I | D^-1] by Gauss-Jordan with partial pivoting
    for (int i = 0; i < h; ++i) {
      for (int j = 0; j < h; ++j) M[(size_t)i * 2 * h + j] = A(h + i, h + j);
      M[(size_t)i * 2 * h + h + i] = 1.0;
    }
    for (int col = 0; col < h; ++col) {
      int piv = col;
      for (int r = col + 1; r < h; ++r)
        if (std::fabs(M[(size_t)r * 2 * h + col]) > std::fabs(M[(size_t)piv * 2 * h + col])) piv = r;
      for (int j = 0; j < 2 * h; ++j) std::swap(M[(size_t)col * 2 * h + j], M[(size_t)piv * 2 * h + j]);
      const double inv = 1.0 / M[(size_t)col * 2 * h + col];
      for (int j = 0; j < 2 * h; ++j) M[(size_t)col * 2 * h + j] *= inv;
      for (int r = 0; r < h; ++r) {
        if (r == col) continue;
        const double f = M[(size_t)r * 2 * h + col];
        for (int j = 0; j < 2 * h; ++j) M[(size_t)r * 2 * h + j] -= f * M[(size_t)col * 2 * h + j];
      }
    }
    for (int i = 0; i < h; ++i) (*Ainv)(i, i) = 1.0 / A(i, i);
    for (int i = 0; i < h; ++i)
      for (int j = 0; j < h; ++j) {
        (*Ainv)(h + i, h + j) = M[(size_t)i * 2 * h + h + j];
        double s = 0.0;  // -(D^-1 C)(i, j) / A(j, j)
        for (int q = 0; q < h; ++q) s += M[(size_t)i * 2 * h + h + q] * A(h + q, j);
        (*Ainv)(h + i, j) = -s / A(j, j);
      }
  }
  // LIN_I:557-573 (this Q is 2x the integral's Hessian; computeCost carries the 1/2)
  static void computeQuadraticCostJacobian(int derivative, double t, SquareMatrix* cost_jacobian) {
    MTG_SHIM_CHECK(derivative < N, "derivative < N");
    MTG_SHIM_CHECK(cost_jacobian != nullptr, "cost_jacobian is null");
    cost_jacobian->resize(N, N);
    for (int a = derivative; a < N; ++a)
      for (int b = derivative; b < N; ++b) {
        const double e = (double)(a + b - 2 * derivative + 1);
        (*cost_jacobian)(a, b) = Polynomial::baseCoefficient(derivative, a) * Polynomial::baseCoefficient(derivative, b) *
                                 std::pow(t, e) * 2.0 / e;
      }
  }

  // LIN_I:500-555: dense views used by the non-linear layer and the ConstraintPacking test
  void getAInverse(MatrixXd* A_inv) const { blockDiagonal(A_inv, true); }
  void getA(MatrixXd* A) const { blockDiagonal(A, false); }
  void getM(MatrixXd* M) const {  // the reordering matrix C: one 1 per row (LIN_I:171-252)
    MTG_SHIM_CHECK(M != nullptr, "M is null");
    M->resize(n_all_constraints_, n_fixed_constraints_ + n_free_constraints_);
    for (size_t r = 0; r < n_all_constraints_; ++r) (*M)(r, columnOfRow(r)) = 1.0;
  }
  void getMpinv(MatrixXd* M_pinv) const {  // row-normalised C^T (every column of C holds 1 or 2 ones)
    MTG_SHIM_CHECK(M_pinv != nullptr, "M_pinv is null");
    const size_t n = n_fixed_constraints_ + n_free_constraints_;
    M_pinv->resize(n, n_all_constraints_);
    std::vector<int> count(n, 0);
    for (size_t r = 0; r < n_all_constraints_; ++r) ++count[columnOfRow(r)];
    for (size_t r = 0; r < n_all_constraints_; ++r) (*M_pinv)(columnOfRow(r), r) = 1.0 / count[columnOfRow(r)];
  }
  void getR(MatrixXd* R) const {  // R = C^T blockdiag(A^-T Q A^-1) C (LIN_I:306-335)
    MTG_SHIM_CHECK(R != nullptr, "R is null");
    const size_t n = n_fixed_constraints_ + n_free_constraints_;
    R->resize(n, n);
    for (size_t i = 0; i < n_segments_; ++i) {
      SquareMatrix A, Ai, Q;
      setupMappingMatrix(segment_times_[i], &A);
      invertMappingMatrix(A, &Ai);
      computeQuadraticCostJacobian(derivative_to_optimize_, segment_times_[i], &Q);
      for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) {
          double hrc = 0.0;  // (A^-T Q A^-1)(r, c)
          for (int a = 0; a < N; ++a) {
            double qa = 0.0;
            for (int b = 0; b < N; ++b) qa += Q(a, b) * Ai(b, c);
            hrc += Ai(a, r) * qa;
          }
          (*R)(columnOfRow(i * N + r), columnOfRow(i * N + c)) += hrc;
        }
    }
  }

  size_t getDimension() const { return dimension_; }
  size_t getNumberSegments() const { return n_segments_; }
  size_t getNumberAllConstraints() const { return n_all_constraints_; }
  size_t getNumberFixedConstraints() const { return n_fixed_constraints_; }
  size_t getNumberFreeConstraints() const { return n_free_constraints_; }

  // ---- batched companion (what a caller with many problems uses; thin veneer over the C ABI).
  // Problems share K, D, N and the canonical pattern; records are AoS (one problem after another).
  //  positions [B][K+1][D], end_derivatives [B][2][N/2-1][D] or nullptr, segment_times [B][K]
  //  -> coefficients [B][K][D][N], cost [B], free [B][D][K-1][N/2-1] (or nullptr), status [B] (or nullptr)
  static void solveLinearBatch(int B, int K, int D, int derivative_to_optimize, const double* positions,
                               const double* end_derivatives, const double* segment_times, double* coefficients,
                               double* cost, double* free_constraints, uint32_t* status) {
    mtg_problem_desc d = runtime::desc(B, K, D, N, derivative_to_optimize);
    runtime::check_rc(mtg_solve_batch(runtime::context(), &d, positions, end_derivatives, segment_times,
                                      coefficients, cost, free_constraints, status, nullptr),
                      "mtg_solve_batch");
  }

 protected:
  // row r of C / of blockdiag(H): segment r / N, local index r % N: < h = (start vertex, k), >= h = (end vertex, k)
  size_t columnOfRow(size_t r) const {
    const size_t h = N / 2;
    const size_t seg = r / N, l = r % N;
    const size_t v = seg + (l >= h ? 1 : 0), k = l % h;
    // fixed constraints first, then free ones, each ordered by (vertex, derivative) (LIN_H:289-296)
    size_t fixed_before = 0, free_before = 0;
    for (size_t q = 0; q < v * h + k; ++q) (mask_[q] ? fixed_before : free_before) += 1;
    return mask_[v * h + k] ? fixed_before : n_fixed_constraints_ + free_before;
  }
  // values [K+1][h][D]: the fixed constraints and, if given, the free derivatives merged in
  void packValues(const std::vector<VectorXd>* free_constraints, std::vector<double>* values) const {
    const int h = N / 2, K = (int)n_segments_, D = (int)dimension_;
    values->assign((size_t)(K + 1) * h * D, 0.0);
    size_t q = 0;
    for (int v = 0; v <= K; ++v)
      for (int k = 0; k < h; ++k) {
        if (mask_[v * h + k]) {
          VectorXd c;
          vertices_[v].getConstraint(k, &c);
          for (int dim = 0; dim < D; ++dim) (*values)[((size_t)v * h + k) * D + dim] = c[dim];
        } else {
          if (free_constraints)
            for (int dim = 0; dim < D; ++dim) (*values)[((size_t)v * h + k) * D + dim] = (*free_constraints)[dim][q];
          ++q;
        }
      }
  }
  void blockDiagonal(MatrixXd* out, bool inverse) const {
    MTG_SHIM_CHECK(out != nullptr, "output is null");
    out->resize(n_all_constraints_, n_all_constraints_);
    for (size_t i = 0; i < n_segments_; ++i) {
      SquareMatrix A, Ai;
      setupMappingMatrix(segment_times_[i], &A);
      if (inverse) invertMappingMatrix(A, &Ai);
      for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) (*out)(i * N + r, i * N + c) = inverse ? Ai(r, c) : A(r, c);
    }
  }
  // positions [K+1][D] and end derivatives [2][h-1][D] for the C ABI
  void packVertices(std::vector<double>* pos, std::vector<double>* endd) const {
    const int h = N / 2, K = (int)n_segments_, D = (int)dimension_;
    for (int v = 0; v <= K; ++v) {
      VectorXd c;
      vertices_[v].getConstraint(derivative_order::POSITION, &c);
      for (int dim = 0; dim < D; ++dim) pos->push_back(c[dim]);
    }
    for (int side = 0; side < 2; ++side)
      for (int m = 1; m < h; ++m) {
        VectorXd c;
        vertices_[side ? K : 0].getConstraint(m, &c);
        for (int dim = 0; dim < D; ++dim) endd->push_back(c[dim]);
      }
  }
  void unpackSegments(const std::vector<double>& coeffs) {
    const int D = (int)dimension_;
    for (size_t i = 0; i < n_segments_; ++i) {
      segments_[i].setTime(segment_times_[i]);
      for (int dim = 0; dim < D; ++dim) {
        VectorXd c(N);
        for (int j = 0; j < N; ++j) c[j] = coeffs[(i * D + dim) * N + j];
        segments_[i][dim] = Polynomial(N, c);
      }
    }
  }

  Vertex::Vector vertices_;
  Segment::Vector segments_;
  std::vector<double> segment_times_;
  std::vector<VectorXd> free_constraints_compact_;
  std::vector<uint8_t> mask_;  // [(K+1)][N/2]: 1 = fixed
  bool canonical_ = true;
  size_t dimension_;
  int derivative_to_optimize_;
  size_t n_vertices_, n_segments_, n_all_constraints_, n_fixed_constraints_, n_free_constraints_;
  double cost_;
};

}  // namespace mav_trajectory_generation
#endif

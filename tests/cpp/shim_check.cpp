// GPU check of the C++ class API (include/mav_tube_trajectory_generation/*.h) in the style of the
// reference's own gtest suite (test/test_polynomial_optimization.cpp): TwoVerticesSetup golden
// coefficients (:707-751), checkPath continuity / constraint satisfaction (:113-195),
// ConstraintPacking (:511-570), ExtremaOfMagnitude (:307-406), evaluateRange semantics.
// Prints one line per check and "SHIM OK" at the end; exits non-zero on the first failure.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include <mav_tube_trajectory_generation/polynomial_optimization_linear.h>

using namespace mav_trajectory_generation;

static int failures = 0;
#define EXPECT(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond);         \
      ++failures;                                                           \
    }                                                                       \
  } while (0)

static void two_vertices_setup() {
  Vertex::Vector vertices(2, Vertex(1));
  vertices[0].makeStartOrEnd(0.0, 4);
  vertices[1].makeStartOrEnd(5.0, 4);
  PolynomialOptimization<10> opt(1);
  EXPECT(opt.setupFromVertices(vertices, std::vector<double>(1, 5.0), derivative_order::SNAP));
  EXPECT(opt.solveLinear());
  Segment::Vector segs;
  opt.getSegments(&segs);
  const double gold[10] = {0, 0, 0, 0, 0, 0.2016, -0.1344, 0.03456, -0.004032, 0.0001792};  // TEST_OPT:741-744
  const VectorXd c = segs[0][0].getCoefficients(0);
  for (int j = 0; j < 10; ++j) EXPECT(std::fabs(c[j] - gold[j]) < 1e-12);
  std::printf("two_vertices_setup: c5..c9 = %.6g %.6g %.6g %.6g %.6g\n", c[5], c[6], c[7], c[8], c[9]);
}

static void random_problem(int D, int K, size_t seed) {
  const VectorXd lo = VectorXd::Constant(D, -10.0), hi = VectorXd::Constant(D, 10.0);
  Vertex::Vector vertices = createRandomVertices(4, K, lo, hi, seed);
  std::vector<double> times = estimateSegmentTimes(vertices, 3.0, 5.0);
  PolynomialOptimization<10> opt(D);
  EXPECT(opt.setupFromVertices(vertices, times, derivative_order::SNAP));
  EXPECT(opt.solveLinear());
  Trajectory traj;
  opt.getTrajectory(&traj);
  EXPECT(traj.K() == K && traj.D() == D && traj.N() == 10);
  // checkPath: fixed constraints are met, derivatives 0..4 are continuous at the interior vertices
  double t_acc = 0.0;
  const Segment::Vector& segs = traj.segments();
  double worst = 0.0;
  for (int i = 0; i < K; ++i) {
    for (int der = 0; der <= 4; ++der) {
      const VectorXd a = segs[i].evaluate(0.0, der), b = segs[i].evaluate(times[i], der);
      VectorXd want;
      if (vertices[i].getConstraint(der, &want))
        for (int d = 0; d < D; ++d) worst = std::fmax(worst, std::fabs(a[d] - want[d]));
      if (vertices[i + 1].getConstraint(der, &want))
        for (int d = 0; d < D; ++d) worst = std::fmax(worst, std::fabs(b[d] - want[d]));
      if (i + 1 < K) {
        const VectorXd n = segs[i + 1].evaluate(0.0, der);
        for (int d = 0; d < D; ++d) worst = std::fmax(worst, std::fabs(b[d] - n[d]));
      }
    }
    t_acc += times[i];
  }
  EXPECT(worst < 1e-6);  // TEST_OPT:116
  EXPECT(std::fabs(traj.getMaxTime() - t_acc) < 1e-12);
  // Trajectory::evaluate: a vertex time belongs to the segment on its right; the end evaluates the last segment
  const VectorXd at_v = traj.evaluate(times[0], 0), seg1 = segs[K > 1 ? 1 : 0].evaluate(K > 1 ? 0.0 : times[0], 0);
  for (int d = 0; d < D; ++d) EXPECT(std::fabs(at_v[d] - seg1[d]) < 1e-12);
  // evaluateRange: serial recurrence, 1000 or 1001 samples for dt = T/1000
  std::vector<VectorXd> samples;
  std::vector<double> st;
  traj.evaluateRange(0.0, traj.getMaxTime(), traj.getMaxTime() / 1000.0, 1, &samples, &st);
  EXPECT(samples.size() == 1000 || samples.size() == 1001);
  EXPECT(st.size() == samples.size() && st[0] == 0.0);
  double acc = 0.0, vmax_sampled = 0.0;
  for (size_t k = 0; k < st.size(); ++k) {
    EXPECT(st[k] == acc);  // bit-exact accumulated time
    acc += traj.getMaxTime() / 1000.0;
    vmax_sampled = std::fmax(vmax_sampled, samples[k].norm());
  }
  // ExtremaOfMagnitude: analytic maximum bounds and matches the sampled one
  std::vector<Extremum> cands;
  const Extremum vmax = opt.computeMaximumOfMagnitude<derivative_order::VELOCITY>(&cands);
  std::vector<int> dims;
  for (int d = 0; d < D; ++d) dims.push_back(d);
  Extremum mn, mx;
  EXPECT(traj.computeMinMaxMagnitude(derivative_order::VELOCITY, dims, &mn, &mx));
  EXPECT(vmax.value >= vmax_sampled - 1e-9 && vmax.value <= vmax_sampled + 0.01);
  EXPECT(std::fabs(vmax.value - mx.value) < 1e-12 && vmax.segment_idx == mx.segment_idx);
  EXPECT((int)cands.size() >= 2 * K + 1);  // per segment [0, T, roots...] + the end of the last segment (LIN_I:455-487)
  // ConstraintPacking: [d_f; d_p] -> p = A^-1 M d -> A p -> M^+ -> [d_f; d_p]; coefficients == segments
  std::vector<VectorXd> d_f, d_p;
  opt.getFixedConstraints(&d_f);
  opt.getFreeConstraints(&d_p);
  MatrixXd Ai, A, M, Mp;
  opt.getAInverse(&Ai);
  opt.getA(&A);
  opt.getM(&M);
  opt.getMpinv(&Mp);
  const size_t nf = opt.getNumberFixedConstraints(), np = opt.getNumberFreeConstraints(), na = opt.getNumberAllConstraints();
  EXPECT(na == (size_t)K * 10 && nf == (size_t)(K - 1 + 10) && np == (size_t)(K - 1) * 4);
  double pack_err = 0.0;
  for (int dim = 0; dim < D; ++dim) {
    std::vector<double> d_all(nf + np), Md(na, 0.0), p(na, 0.0), Ap(na, 0.0), back(nf + np, 0.0);
    for (size_t q = 0; q < nf; ++q) d_all[q] = d_f[dim][q];
    for (size_t q = 0; q < np; ++q) d_all[nf + q] = d_p[dim][q];
    for (size_t r = 0; r < na; ++r)
      for (size_t c = 0; c < nf + np; ++c) Md[r] += M(r, c) * d_all[c];
    for (size_t r = 0; r < na; ++r)
      for (size_t c = 0; c < na; ++c) p[r] += Ai(r, c) * Md[c];
    for (size_t r = 0; r < na; ++r)
      for (size_t c = 0; c < na; ++c) Ap[r] += A(r, c) * p[c];
    for (size_t r = 0; r < nf + np; ++r)
      for (size_t c = 0; c < na; ++c) back[r] += Mp(r, c) * Ap[c];
    for (size_t q = 0; q < nf + np; ++q) pack_err = std::fmax(pack_err, std::fabs(back[q] - d_all[q]));
    for (int i = 0; i < K; ++i) {
      const VectorXd c = segs[i][dim].getCoefficients(0);
      double scale = 0.0, err = 0.0;
      for (int j = 0; j < 10; ++j) {
        scale = std::fmax(scale, std::fabs(c[j]));
        err = std::fmax(err, std::fabs(c[j] - p[i * 10 + j]));
      }
      EXPECT(err <= 1e-6 * std::fmax(scale, 1.0));  // TEST_OPT:565-568 tolerance
    }
  }
  EXPECT(pack_err < 1e-6);
  // cost: 0.5 d^T R d with the host-assembled R (LIN_I:113-130 vs NL_I:1585-1588)
  MatrixXd R;
  opt.getR(&R);
  double J = 0.0;
  for (int dim = 0; dim < D; ++dim) {
    std::vector<double> d_all(nf + np);
    for (size_t q = 0; q < nf; ++q) d_all[q] = d_f[dim][q];
    for (size_t q = 0; q < np; ++q) d_all[nf + q] = d_p[dim][q];
    for (size_t r = 0; r < nf + np; ++r)
      for (size_t c = 0; c < nf + np; ++c) J += d_all[r] * R(r, c) * d_all[c];
  }
  EXPECT(std::fabs(0.5 * J - opt.computeCost()) <= 1e-6 * opt.computeCost());
  // setFreeConstraints round trip and optimality of d_p
  const double cost0 = opt.computeCost();
  std::vector<VectorXd> bumped = d_p;
  if (np) bumped[0][0] *= 1.1;
  opt.setFreeConstraints(bumped);
  EXPECT(np == 0 || opt.computeCost() > cost0);
  opt.setFreeConstraints(d_p);
  EXPECT(std::fabs(opt.computeCost() - cost0) <= 1e-13 * cost0);
  std::printf("random_problem D=%d K=%d seed=%zu: cost %.9g, checkPath %.2e, packing %.2e, max|v| %.6g (sampled %.6g)\n",
              D, K, seed, cost0, worst, pack_err, vmax.value, vmax_sampled);
}

int main() {
  two_vertices_setup();
  random_problem(3, 10, 105);  // segment_10_dim_3, TEST_OPT:786-792 = BASELINE configs[0]
  random_problem(1, 10, 102);
  random_problem(3, 1, 104);
  random_problem(3, 50, 106);
  // a general constraint pattern (interior velocity fixed, free goal derivatives) takes the generic solver:
  // constraints met, C^4 continuity, optimality of d_p, ConstraintPacking identities
  {
    Vertex::Vector v = createRandomVertices(4, 6, VectorXd::Constant(3, -5.0), VectorXd::Constant(3, 5.0), 7);
    v[2].addConstraint(derivative_order::VELOCITY, 0.5);
    v[4].addConstraint(derivative_order::VELOCITY, VectorXd::Constant(3, -0.25));
    for (int k = 2; k <= 4; ++k) v[6].removeConstraint(k);
    std::vector<double> times = estimateSegmentTimes(v, 3.0, 5.0);
    PolynomialOptimization<10> opt(3);
    EXPECT(opt.setupFromVertices(v, times));
    EXPECT(opt.getNumberFixedConstraints() == 5 + 2 + 5 + 2 && opt.getNumberFreeConstraints() == 35 - 14);  // start, goal, 5 positions, 2 velocities
    EXPECT(opt.solveLinear());
    Segment::Vector segs;
    opt.getSegments(&segs);
    double worst = 0.0;
    for (int i = 0; i < 6; ++i)
      for (int der = 0; der <= 4; ++der) {
        const VectorXd a = segs[i].evaluate(0.0, der), b = segs[i].evaluate(times[i], der);
        VectorXd want;
        if (v[i].getConstraint(der, &want))
          for (int d = 0; d < 3; ++d) worst = std::fmax(worst, std::fabs(a[d] - want[d]));
        if (v[i + 1].getConstraint(der, &want))
          for (int d = 0; d < 3; ++d) worst = std::fmax(worst, std::fabs(b[d] - want[d]));
        if (i + 1 < 6) {
          const VectorXd n = segs[i + 1].evaluate(0.0, der);
          for (int d = 0; d < 3; ++d) worst = std::fmax(worst, std::fabs(b[d] - n[d]));
        }
      }
    EXPECT(worst < 1e-6);
    std::vector<VectorXd> d_p;
    opt.getFreeConstraints(&d_p);
    const double cost0 = opt.computeCost();
    std::vector<VectorXd> bumped = d_p;
    bumped[1][3] += 0.05;
    opt.setFreeConstraints(bumped);
    EXPECT(opt.computeCost() > cost0);
    opt.setFreeConstraints(d_p);
    EXPECT(std::fabs(opt.computeCost() - cost0) <= 1e-12 * cost0);
    // 0.5 d^T R d with the host-assembled R and the general column map
    std::vector<VectorXd> d_f;
    opt.getFixedConstraints(&d_f);
    MatrixXd R;
    opt.getR(&R);
    const size_t nf = opt.getNumberFixedConstraints(), np = opt.getNumberFreeConstraints();
    double J = 0.0;
    for (int dim = 0; dim < 3; ++dim) {
      std::vector<double> d_all(nf + np);
      for (size_t q = 0; q < nf; ++q) d_all[q] = d_f[dim][q];
      for (size_t q = 0; q < np; ++q) d_all[nf + q] = d_p[dim][q];
      for (size_t r = 0; r < nf + np; ++r)
        for (size_t c = 0; c < nf + np; ++c) J += d_all[r] * R(r, c) * d_all[c];
    }
    EXPECT(std::fabs(0.5 * J - cost0) <= 1e-6 * cost0);
    std::printf("general pattern: cost %.9g, checkPath %.2e, n_fixed %zu, n_free %zu\n", cost0, worst, nf, np);
  }
  // candidate lists (LIN_I:396-487, segment.cpp:82-158), Polynomial::computeMinMax (polynomial.cpp:99-114, in the
  // spirit of PolynomialTest.FindMinMax TEST_POLY:81-137) and computeCost after updateSegmentTimes (LIN_I:113-130)
  {
    Vertex::Vector v = createRandomVertices(4, 5, VectorXd::Constant(3, -10.0), VectorXd::Constant(3, 10.0), 109);
    std::vector<double> times = estimateSegmentTimes(v, 3.0, 5.0);
    PolynomialOptimization<10> opt(3);
    EXPECT(opt.setupFromVertices(v, times));
    EXPECT(opt.solveLinear());
    std::vector<Extremum> cand;
    const Extremum vmax = opt.computeMaximumOfMagnitude(derivative_order::VELOCITY, &cand);
    EXPECT(cand.size() >= 2 * 5 + 1);
    double best = 0.0;
    for (const Extremum& e : cand) best = std::fmax(best, e.value);
    EXPECT(best == vmax.value);
    Segment::Vector segs;
    opt.getSegments(&segs);
    std::vector<double> ct;
    EXPECT(PolynomialOptimization<10>::computeSegmentMaximumMagnitudeCandidates(derivative_order::VELOCITY, segs[2], 0.0,
                                                                                segs[2].getTime(), &ct));
    EXPECT(ct.size() >= 2 && ct[0] == 0.0 && ct[1] == segs[2].getTime());
    for (size_t q = 2; q < ct.size(); ++q) {   // every root is a stationary point of |v|^2: v . a = 0
      const VectorXd vel = segs[2].evaluate(ct[q], 1), acc = segs[2].evaluate(ct[q], 2);
      double dot = 0.0, nv = 0.0, na = 0.0;
      for (int d = 0; d < 3; ++d) { dot += vel[d] * acc[d]; nv += vel[d] * vel[d]; na += acc[d] * acc[d]; }
      EXPECT(std::fabs(dot) <= 1e-9 * std::sqrt(nv * na) + 1e-12);
    }
    std::vector<double> cs;
    PolynomialOptimization<10>::computeSegmentMaximumMagnitudeCandidatesBySampling<derivative_order::VELOCITY>(
        segs[2], 0.0, segs[2].getTime(), 0.01, &cs);
    for (double t : cs) {   // every sampled candidate sits next to an analytic one
      double d = 1e9;
      for (double a : ct) d = std::fmin(d, std::fabs(a - t));
      EXPECT(d <= 0.02);
    }
    // one polynomial: extrema of p'(t) on a sub-interval against dense sampling
    std::pair<double, double> mn, mx;
    EXPECT(segs[1][0].computeMinMax(0.2, segs[1].getTime() * 0.9, 1, &mn, &mx));
    double smin = 1e300, smax = -1e300;
    for (double t = 0.2; t <= segs[1].getTime() * 0.9; t += 0.001) {
      const double val = segs[1][0].evaluate(t, 1);
      smin = std::fmin(smin, val);
      smax = std::fmax(smax, val);
    }
    EXPECT(mx.second >= smax - 1e-12 && mx.second <= smax + 1e-4 && mn.second <= smin + 1e-12 && mn.second >= smin - 1e-4);
    // new segment times without a new solve: the cost is 0.5 sum c^T Q(T_new) c of the OLD coefficients
    const double cost0 = opt.computeCost();
    std::vector<double> t2 = times;
    for (double& x : t2) x *= 1.5;
    opt.updateSegmentTimes(t2);
    double want = 0.0;
    for (int i = 0; i < 5; ++i) {
      MatrixXd Q;
      PolynomialOptimization<10>::computeQuadraticCostJacobian(derivative_order::SNAP, t2[i], &Q);
      for (int d = 0; d < 3; ++d) {
        const VectorXd c = segs[i][d].getCoefficients(0);
        for (int a = 0; a < 10; ++a)
          for (int b = 0; b < 10; ++b) want += 0.5 * c[a] * Q(a, b) * c[b];
      }
    }
    EXPECT(std::fabs(opt.computeCost() - want) <= 1e-9 * want && std::fabs(want - cost0) > 1e-3 * cost0);
    std::printf("candidates: %zu over 5 segments, max|v| %.6g; computeCost after updateSegmentTimes %.9g (was %.9g)\n",
                cand.size(), vmax.value, opt.computeCost(), cost0);
  }
  if (failures) {
    std::printf("SHIM FAILED: %d check(s)\n", failures);
    return 1;
  }
  std::printf("SHIM OK\n");
  return 0;
}

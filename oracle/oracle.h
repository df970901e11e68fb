/* TEST INFRASTRUCTURE — CPU oracle for the batched min-snap hot path.
 *
 * This is a plain-C++ (no Eigen, no glog) single-threaded RESTATEMENT of the
 * reference algorithm of NilsFunk/mav_tube_trajectory_generation, written from
 * the cited lines. It is the checker, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load liboracle.so. The product path (libmtg_cuda.so) never links it.
 *
 * Parity pinning:
 *   - pinned by reference golden vectors / KATs: P2-P5, P7, P8c (TwoVerticesSetup
 *     Matlab coefficients TEST_OPT:741-744, AMatrixInversion TEST_OPT:695-705,
 *     Convolution TEST_POLY:68-79), R1 core (the reference's own rpoly_ak1.cpp is
 *     compiled verbatim into oracle/_ref/librpoly_ref.so and linked here).
 *   - pinned only through invariants (checkPath 1e-6, checkCost 10 %,
 *     ConstraintPacking 1e-6) + an 80-digit mpmath solve: the d_p solve P8b
 *     (reference uses Eigen::SparseQR, third-party, unvendored, version unpinned).
 *   - PARITY UNPINNED: T1 (sampled tube predicate; the reference only feeds the
 *     geometry to MOSEK, no test pins it).
 *
 * Citation tags (relative to /root/reference): LIN_I = include/.../impl/
 * polynomial_optimization_linear_impl.h, LIN_H, POLY_H, POLY_C = src/polynomial.cpp,
 * SEG_C, TRAJ_C, VTX_C, RPOLY_C, QC_I, NL_I as in SURVEY.md.
 *
 * Layouts: matrices row-major; coefficients [K][D][N] (segment, dimension,
 * increasing power); vertex constraints: mask[(K+1)][N/2] (1 = fixed),
 * values[(K+1)][N/2][D].
 */
#ifndef MTG_ORACLE_H_
#define MTG_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTGO_MAX_N 12
#define MTGO_MAX_CONV 22 /* POLY_H:45-48 */

/* P2  POLY_C:145-161,200-201: 22x22 table B(n,i) = i!/(i-n)! */
void mtgo_base_coefficients(double* out22x22);
/* P3  LIN_I:557-573 */
void mtgo_quadratic_cost_jacobian(int N, int derivative, double t, double* Q);
/* POLY_H:201-228 */
void mtgo_base_coeffs_with_time(int N, int derivative, double t, double* out);
/* P4  LIN_I:101-111 */
void mtgo_setup_mapping_matrix(int N, double t, double* A);
/* P5  LIN_I:132-169 (Schur block inverse, D block by partial-pivot LU) */
void mtgo_invert_mapping_matrix(int N, const double* A, double* Ainv);
/* dense partial-pivot inverse (stands in for Eigen's A.inverse(), TEST_OPT:701) */
int mtgo_general_inverse(int n, const double* M, double* Minv);

/* G1  VTX_C:27-82 (std::mt19937 + uniform_real_distribution, 0.2 m rejection).
 * mask/values use half_n rows per vertex. Returns number of vertices. */
int mtgo_create_random_vertices(int maximum_derivative, int n_segments, int D,
                                const double* pos_min, const double* pos_max,
                                uint64_t seed, int half_n, uint8_t* mask,
                                double* values);
/* VTX_C:252-269 / 233-250,271-287; positions [(K+1)][D] */
void mtgo_estimate_segment_times_nfabian(int K, int D, const double* positions,
                                         double v_max, double a_max,
                                         double magic_fabian_constant,
                                         double* times);
void mtgo_estimate_segment_times_velocity_ramp(int K, int D,
                                               const double* positions,
                                               double v_max, double a_max,
                                               double* times);

/* P1,P6,P7,P8a-d. solver: 0 = dense Householder QR (closest to the reference's
 * SparseQR, LIN_I:364-374), 1 = dense Cholesky. Optional outputs may be NULL.
 * counts = {n_all, n_fixed, n_free}. col_of_row[n_all] is the reordering matrix
 * C (one 1 per row). R_out is (n_fixed+n_free)^2. Returns 0 on success,
 * negative on invalid input (the reference CHECK-aborts there). */
int mtgo_solve(int N, int D, int K, int derivative, const double* times,
               const uint8_t* mask, const double* values, int solver,
               double* coeffs, double* cost, int* counts, double* d_f,
               double* d_p, double* R_out, int* col_of_row);

/* Batched form over the CANONICAL pattern (createRandomVertices: rest-to-rest
 * ends, position-only interior) for CPU-baseline timing. positions [B][K+1][D],
 * times [B][K]; coeffs [B][K][D][N]; n_threads>1 uses OpenMP over the batch. */
int mtgo_solve_canonical_batch(int B, int N, int D, int K, int derivative,
                               const double* positions, const double* times,
                               int solver, int n_threads, double* coeffs,
                               double* cost);

/* LIN_I:489-498 + 254-275: coefficients from given d_p (setFreeConstraints). */
int mtgo_coeffs_from_free_constraints(int N, int D, int K, const double* times,
                                      const uint8_t* mask, const double* values,
                                      const double* d_p, double* coeffs);

/* P9  NL_I:1537-1606 (J_d, no 1/2) and NL_I:2495-2657 (central / forward FD,
 * floor 0.1, d_p held fixed). J_plus/J_minus [K] (J_minus untouched if
 * !central); grad_d[K] is dJd/dT only (weights applied by the caller). */
int mtgo_cost_time_fd(int N, int D, int K, int derivative, const double* times,
                      const uint8_t* mask, const double* values,
                      const double* d_p, double increment_time, int central,
                      double* J_nominal, double* J_plus, double* J_minus,
                      double* grad_d);

/* E1  POLY_H:136-149 */
double mtgo_poly_evaluate(int N, const double* c, double t, int derivative);
/* POLY_H:99-113 */
void mtgo_poly_derivative_coefficients(int N, const double* c, int derivative,
                                       double* out);
/* POLY_C:163-181 */
void mtgo_convolve(const double* data, int n_data, const double* kernel,
                   int n_kernel, double* out);
/* E3  TRAJ_C:41-72. Returns segment index used, or -1 if t out of range
 * (out zero-filled like the reference). */
int mtgo_traj_evaluate(int N, int D, int K, const double* coeffs,
                       const double* times, double t, int derivative,
                       double* out);
/* E4  TRAJ_C:74-134. Returns the number of samples written (<= cap), or -1 if
 * t_start is out of range (reference returns empty; t_start == max_time is UB
 * in the reference and defined as an error here). */
int mtgo_traj_evaluate_range(int N, int D, int K, const double* coeffs,
                             const double* times, double t_start, double t_end,
                             double dt, int derivative, int cap, double* out,
                             double* sampling_times, int32_t* segment_idx);

/* R1 wrapper RPOLY_C:57-117 restated; core = reference rpoly (oracle/_ref).
 * Returns the wrapper's bool (1/0), or -2 if built without the reference. */
int mtgo_find_roots_jenkins_traub(const double* coeffs_increasing, int n,
                                  double* roots_real, double* roots_imag,
                                  int* n_roots);
/* POLY_C:32-63 */
int mtgo_select_min_max_candidates_from_roots(double t_start, double t_end,
                                              const double* re, const double* im,
                                              int n_roots, double* candidates);
/* POLY_C:102-143: min/max of one polynomial derivative on [t0,t1] */
int mtgo_poly_compute_min_max(int N, const double* c, double t_start,
                              double t_end, int derivative, double* min_t,
                              double* min_v, double* max_t, double* max_v);
/* E6  SEG_C:82-133: candidate times for one segment; seg_coeffs [D][N];
 * dims lists the dimensions used. Returns count (>=2) or -1 on failure. */
int mtgo_segment_candidate_times(int N, int D, const double* seg_coeffs,
                                 int derivative, double t_start, double t_end,
                                 const int* dims, int n_dims,
                                 double* candidate_times, int cap);
/* TRAJ_C:184-220 */
int mtgo_traj_min_max_magnitude(int N, int D, int K, const double* coeffs,
                                const double* times, int derivative,
                                const int* dims, int n_dims, double* min_time,
                                double* min_value, int* min_seg,
                                double* max_time, double* max_value,
                                int* max_seg);
/* LIN_I:455-487 */
int mtgo_opt_max_magnitude(int N, int D, int K, const double* coeffs,
                           const double* times, int derivative,
                           double* max_time, double* max_value, int* max_seg);
/* test_utils.h:43-54 getMaximumMagnitude (dt sampling via E3) */
double mtgo_sampled_maximum_magnitude(int N, int D, int K, const double* coeffs,
                                      const double* times, int derivative,
                                      double dt);
/* test_utils.h:56-64 computeCostNumeric */
double mtgo_cost_numeric(int N, int D, int K, const double* coeffs,
                         const double* times, int derivative, double dt);

/* T1  QC_I:357-474 geometry. positions [(K+1)][3], radii [K][2]
 * (first = tube radius, second = end sphere / cap radius). Output per segment:
 * geom[K][24] = {A(9), b(3), n(3), p_start(3), p_end(3), r_tube, r_sphere,
 * pad} (see oracle.cpp). */
void mtgo_tube_geometry(int K, const double* positions, const double* radii,
                        double* geom);
/* sampled predicate for one 3-D point in segment seg. bit0 in_tube (incl. caps),
 * bit1 in_sphere (end vertex sphere). PARITY UNPINNED. */
int mtgo_tube_flags(const double* geom_seg, const double* vertex_end,
                    const double* x);

/* Fused sweep used as the checker of mtg_feasibility_batch: evaluateRange at
 * position/velocity/acceleration on the same sampling recurrence, flags bit0
 * |v|<=v_max, bit1 |a|<=a_max, bit2 in_tube. Returns n samples. */
int mtgo_feasibility_sweep(int N, int K, const double* coeffs,
                           const double* times, const double* positions,
                           const double* radii, double v_max, double a_max,
                           double t_start, double t_end, double dt, int cap,
                           double* pos_out, uint8_t* flags, double* max_v,
                           double* max_a);

/* N2  QC_I:267-319: inverse control-point mapping matrix of one segment (N x N, row-major):
 * Bezier control points = B_inv [d(v_i, 0..h-1); d(v_i+1, 0..h-1)], incl. the reference's 1e-5 zeroing. */
int mtgo_inverse_control_point_mapping(int N, double T, double* B_inv);
/* N2  QC_I:321-474: control points [K][N][D] of a trajectory given the full endpoint derivatives of every
 * vertex [K+1][h][D] (= C [d_f; d_p]) and, for D = 3 with positions/radii, the VALUES of the reference's
 * constraints on them (feasible <=> value <= 0): tube / cap_start / cap_end [K][N-2] on control points
 * 1..N-2, sphere [K] on the last control point (-inf for the last segment: not constrained there).
 * No reference test pins these (the reference hands them to MOSEK); they are pinned here by two
 * mathematical properties (tests/test_oracle.py): the control points reproduce the polynomial in the
 * Bernstein basis, and control points inside the convex tube-and-caps region imply every sample inside. */
int mtgo_control_point_constraints(int N, int K, int D, const double* derivatives, const double* times,
                                   const double* positions, const double* radii, double* control_points,
                                   double* tube, double* cap_start, double* cap_end, double* sphere);

/* N1  NL_I:1537-1606: J_d (no 1/2) and grad_{d_p} = 2 R_pf d_f + 2 R_pp d_p [D][n_free] from the dense R. */
int mtgo_cost_gradient_derivative(int N, int D, int K, int derivative, const double* times, const uint8_t* mask,
                                  const double* values, const double* d_p, double* J_d, double* grad);
/* N1  NL_I:2735-2766, 2365-2490: soft-constraint cost and its finite-difference gradient wrt d_p [D][n_free]. */
int mtgo_soft_constraint_gradient(int N, int D, int K, int derivative, const double* times, const uint8_t* mask,
                                  const double* values, const double* d_p, int n_con, const int* ders,
                                  const double* limits, double weight, double max_cost, double increment,
                                  int central, double* J_sc, double* grad);

/* N4  NL_I:1608-1780, 1783-1917, 2659-2684: collision potential along the trajectory against a dense grid of
 * distances [size0][size1][size2] (metres; element [0][0][0] is voxel `origin`), canonical constraint pattern.
 * grad [3][(K-1)(N/2-1)] or NULL. */
int mtgo_collision_cost(int N, int K, const double* coeffs, const double* times, const double* grid,
                        const int* size, const int* origin, double res, const double* min_bound,
                        const double* max_bound, double dt, double epsilon, double robot_radius, double multiplier,
                        double* J_c, double* grad, int* in_collision, int* n_checks);

/* N3  NL_I:2907-3003: the sampled dump [t, pos, vel, acc, jerk, snap, tm] as a max_rows x (5 D + 2) matrix. */
int mtgo_sample_dump(int N, int D, int K, const double* coeffs, const double* times, double dt, int max_rows,
                     double* rows);

int mtgo_has_reference_rpoly(void);

#ifdef __cplusplus
}
#endif
#endif /* MTG_ORACLE_H_ */

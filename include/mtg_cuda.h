/* mtg_cuda.h — C ABI of libmtg_cuda.so, the B200 (sm_100a) batched implementation
 * of the unconstrained polynomial-trajectory hot path of
 * NilsFunk/mav_tube_trajectory_generation.
 *
 * The reference has NO plugin/FFI boundary: its boundary is the C++ class API
 * (Vertex / Polynomial / Segment / Trajectory / PolynomialOptimization<N>).
 * the headers under include/mav_tube_trajectory_generation/ in this repo re-creates that class
 * API on top of these entry points (B = 1 per object); batched callers (bench,
 * sweeps) call them directly. Every entry point cites the reference interface
 * it replaces (paths relative to the reference root;
 * LIN_I = include/mav_tube_trajectory_generation/impl/polynomial_optimization_linear_impl.h).
 *
 * Conventions
 *  - plain C: pointers + sizes, no C++/torch types. `stream` is a cudaStream_t
 *    passed as void* (NULL = default stream).
 *  - return value: 0 = MTG_OK, negative = call-level error (bad argument, CUDA
 *    error; text via mtg_last_error). Reference programmer-error CHECKs
 *    (process abort there) become negative return codes here.
 *  - per-item problems never abort: they set bits in status[b] (MTG_ST_*).
 *  - there is NO CPU fallback: without a CUDA device mtg_create fails.
 *  - memory: desc.memory says whether ALL data pointers of the call are device
 *    pointers (MTG_MEM_DEVICE; nothing is copied, the call only enqueues work
 *    on `stream`) or host pointers (MTG_MEM_HOST; the library stages H2D/D2H
 *    itself in pipelined chunks and returns when the outputs are in place).
 *  - layout: see MTG_LAYOUT_* below. Tensors are documented by the element
 *    order inside one trajectory's record, e.g. coeffs "[K][D][N]" means
 *    elem = (k*D + d)*N + n. With B = 1 both layouts coincide with the
 *    reference's natural per-object order.
 */
#ifndef MTG_CUDA_H_
#define MTG_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTG_ABI_VERSION 2

#define MTG_OK 0
#define MTG_ERR_INVALID_ARGUMENT (-1) /* reference: glog CHECK abort            */
#define MTG_ERR_CUDA (-2)             /* CUDA runtime error, see mtg_last_error  */
#define MTG_ERR_NO_DEVICE (-3)        /* no CUDA device: the library has no CPU path */
#define MTG_ERR_UNSUPPORTED (-4)
#define MTG_ERR_NCCL (-5)

/* per-item status bits */
#define MTG_ST_OK 0u
#define MTG_ST_BAD_TIME 1u      /* a segment time <= 0 or non-finite (LIN_I:296 CHECK_GT) */
#define MTG_ST_NOT_SPD 2u       /* R_pp pivot <= 0: ill-scaled problem                      */
#define MTG_ST_OUT_OF_RANGE 4u  /* evaluation time outside the trajectory (TRAJ_C:58-61,104-107) */
#define MTG_ST_TRUNCATED 8u     /* more samples than max_samples                            */
#define MTG_ST_NO_CONVERGENCE 16u /* root iteration hit its cap (RPOLY_C:372-377)          */

#define MTG_MEM_DEVICE 0
#define MTG_MEM_HOST 1

/* Every per-trajectory tensor is a batch of B fixed-size RECORDS; a record's
 * element order is given per tensor below ("elem"). desc.layout places the batch:
 *   MTG_LAYOUT_SOA  element-major, batch innermost:  x[elem * B + b]
 *   MTG_LAYOUT_AOS  record-contiguous:               x[b * record_len + elem]
 * AOS is the reference's natural per-object order (Segment -> Polynomial ->
 * coefficients) and is what the C++ shim uses; SOA gives fully coalesced accesses
 * for one-thread-per-trajectory kernels. */
#define MTG_LAYOUT_SOA 0
#define MTG_LAYOUT_AOS 1

#define MTG_MAX_N 12 /* Polynomial::kMaxN, polynomial.h:45 */

typedef struct mtg_ctx mtg_ctx;

/* Shape of one batch. Mirrors the reference's "config" (ctor/setup arguments):
 * PolynomialOptimization<N>(dimension) + setupFromVertices(vertices, times,
 * derivative_to_optimize)  [polynomial_optimization_linear.h:57,67-69]. */
typedef struct mtg_problem_desc {
  int32_t B;                      /* trajectories in the batch (leading dimension) */
  int32_t K;                      /* segments per trajectory (vertices = K+1)      */
  int32_t D;                      /* dimensions (1..4)                             */
  int32_t N;                      /* coefficients per polynomial, even, <= 12      */
  int32_t derivative_to_optimize; /* 0 .. N/2-1                                    */
  int32_t memory;                 /* MTG_MEM_DEVICE / MTG_MEM_HOST                 */
  int32_t layout;                 /* MTG_LAYOUT_SOA / MTG_LAYOUT_AOS               */
} mtg_problem_desc;

/* ------------------------------------------------------------------ context */
int mtg_abi_version(void);
int mtg_create(int device, mtg_ctx** ctx);
void mtg_destroy(mtg_ctx* ctx);
const char* mtg_last_error(const mtg_ctx* ctx);
/* number of kernels this context has launched so far (bench "gpu_launches") */
uint64_t mtg_launch_count(const mtg_ctx* ctx);
int mtg_sync(mtg_ctx* ctx, void* stream);

/* Measured fp64 roof of this device: a register-only DFMA micro-benchmark (8 independent chains per
 * thread, 8 CTAs x 256 threads per SM), timed with CUDA events on `stream`. tflops = 2 x DFMA / s.
 * (Diagnostics for the roofline report; no reference counterpart.) */
int mtg_probe_fp64_fma(mtg_ctx* ctx, int reps, double* tflops, double* ms_per_launch, void* stream);

/* Constant tables of (N, derivative): H1 = A(1)^-T Q(1) A(1)^-1 and A(1)^-1,
 * row-major N x N, computed in binary128 on the host and rounded once. These
 * replace computeQuadraticCostJacobian / setupMappingMatrix /
 * invertMappingMatrix [LIN_I:557-573, 101-111, 132-169] through
 * H(T) = T^(1-2d) S H1 S,  A(T)^-1 = diag(T^-j) A(1)^-1 S,  S = diag(T^alpha). */
int mtg_get_tables(int N, int derivative, double* H1, double* Ainv1);

/* ------------------------------------------------------ P1..P8: linear solve
 * Replaces, per trajectory: setupFromVertices + updateSegmentTimes +
 * setupConstraintReorderingMatrix + constructR + solveLinear +
 * updateSegmentsFromCompactConstraints + computeCost
 * [LIN_I:46-99, 277-304, 171-252, 306-335, 337-379, 254-275, 113-130]
 * for the constraint pattern createRandomVertices produces [src/vertex.cpp:27-82]:
 * first and last vertex fix derivatives 0..N/2-1, interior vertices fix position.
 *
 *  (record element orders; batch placement per desc.layout)
 *  positions        [K+1][D]         in
 *  end_derivatives  [2][N/2-1][D]    in, derivatives 1..N/2-1 at the first ([0]) and
 *                                    last ([1]) vertex; NULL = all zero (makeStartOrEnd,
 *                                    src/vertex.cpp:147-153)
 *  seg_times        [K]              in
 *  coeffs           [K][D][N]        out or NULL, increasing powers (polynomial.h:35-36); NULL = a
 *                                    cost-only solve (candidate sweeps: 8 instead of 2,408 bytes out)
 *  cost             [B]              out or NULL, computeCost() = 0.5 sum c^T Q c
 *  free_constraints [D][K-1][N/2-1]  out or NULL, d_p exactly as getFreeConstraints
 *                                    returns it: per dimension, vertex-major,
 *                                    derivative 1.. within a vertex (LIN_H:289-296)
 *  status           [B] uint32       out or NULL                                   */
int mtg_solve_batch(mtg_ctx* ctx, const mtg_problem_desc* desc,
                    const double* positions, const double* end_derivatives,
                    const double* seg_times, double* coeffs, double* cost,
                    double* free_constraints, uint32_t* status, void* stream);

/* mtg_solve_generic_batch: the same solve for an ARBITRARY constraint pattern shared by the batch
 * — the general setupConstraintReorderingMatrix [LIN_I:171-252]: any subset of the derivatives
 * 0..N/2-1 may be fixed at any vertex.
 *  mask   [(K+1)][N/2] uint8, HOST pointer (small, shared by all B trajectories): 1 = fixed
 *  values [K+1][N/2][D]  in, records: the fixed values (entries of free derivatives are ignored)
 *  free_constraints [D][n_free]  out or NULL: d_p as getFreeConstraints orders it (per dimension,
 *                    by vertex then derivative, LIN_H:289-296); n_free = number of zeros in mask
 * status: MTG_ST_NOT_SPD when a pivot of R_pp is not positive or cancels below 1e-11 of its
 * pre-elimination diagonal (an under-determined pattern, e.g. no position fixed anywhere with a
 * derivative cost; deficiency hidden by rounding is not detectable). Other tensors as mtg_solve_batch. */
int mtg_solve_generic_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const uint8_t* mask,
                            const double* values, const double* seg_times, double* coeffs, double* cost,
                            double* free_constraints, uint32_t* status, void* stream);

/* mtg_set_free_constraints_batch: setFreeConstraints(d_p) + updateSegmentsFromCompactConstraints
 * + computeCost [LIN_I:489-498, 254-275, 113-130]: coefficients and cost of trajectories whose free
 * derivatives are GIVEN (the optimiser-driven call of the non-linear layer, NL_I:1309-1310).
 * Tensors as mtg_solve_batch, with free_constraints [D][K-1][N/2-1] an input. */
int mtg_set_free_constraints_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* positions,
                                   const double* end_derivatives, const double* seg_times,
                                   const double* free_constraints, double* coeffs, double* cost,
                                   uint32_t* status, void* stream);

/* mtg_coeffs_from_derivatives_batch: updateSegmentsFromCompactConstraints + computeCost
 * [LIN_I:254-275, 113-130] for ANY constraint pattern: the caller merges fixed and free entries into
 * the full endpoint derivatives of every vertex (what C [d_f; d_p] is), derivatives [K+1][N/2][D];
 * coeffs = A(T)^-1 [d(v_i); d(v_i+1)] per segment, cost = 0.5 sum c^T Q c. */
int mtg_coeffs_from_derivatives_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* derivatives,
                                      const double* seg_times, double* coeffs, double* cost,
                                      uint32_t* status, void* stream);

/* ------------------------------------- P9: finite-difference time perturbations
 * Replaces the per-segment perturbation loop of the reference's non-linear layer,
 * getCostAndGradientTime [impl/polynomial_optimization_nonlinear_impl.h:2495-2584,
 * central differences] / getCostAndGradientTimeSimple [:2586-2657, forward] with
 * getCostAndGradientDerivative [:1537-1606] as the cost: for every segment n
 *   T(+-)[n] = T[n] <= 0.1 ? 0.1 : T[n] +- increment_time      (:2527-2530, :2547-2550)
 *   J_d(T(+-)) = sum_dim [d_f; d_p]^T R(T(+-)) [d_f; d_p]       (no 1/2; d_p HELD FIXED)
 * i.e. 1 + K (forward) or 1 + 2K (central) rebuilds of every segment's Q, A^-1 and R
 * per optimiser evaluation in the reference, one kernel here.
 *  positions, end_derivatives, seg_times   as mtg_solve_batch
 *  free_constraints [D][K-1][N/2-1]        in, d_p as mtg_solve_batch returns it
 *  J_nominal [B]   out or NULL, J_d(T)
 *  J_plus    [K]   out or NULL, J_d with segment n lengthened
 *  J_minus   [K]   out or NULL, J_d with segment n shortened (central only)
 *  grad      [K]   out or NULL, dJ_d/dT_n: (J_plus - J_minus)/(2 inc) or (J_plus - J_nominal)/inc,
 *                  formed from the perturbed segment's own term (the other K-1 cancel exactly)
 * status: MTG_ST_BAD_TIME also when a shortened time T[n] - increment_time is not positive (central
 * mode, 0.1 < T[n] <= increment_time; the reference aborts in updateSegmentTimes, LIN_I:296):
 * J_minus[n] and grad[n] are then NaN.
 * The caller applies the weights (w_d, w_t ...) of NL_I:2573. */
int mtg_cost_time_fd_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* positions,
                           const double* end_derivatives, const double* seg_times,
                           const double* free_constraints, double increment_time, int central,
                           double* J_nominal, double* J_plus, double* J_minus, double* grad,
                           uint32_t* status, void* stream);

/* --------------------------------------------- E1..E4: sampled evaluation
 * All take the solve's output layout: coeffs [K][D][N], seg_times [K] records.
 *
 * mtg_max_time_batch: Trajectory::getMaxTime() [trajectory.h:63-72,84]: the
 * segment times summed in segment order. max_time [B]. */
int mtg_max_time_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* seg_times,
                       double* max_time, void* stream);

/* mtg_eval_range_batch: Trajectory::evaluateRange(t_start, t_end, dt, derivative,
 * &result, &sampling_times) [src/trajectory.cpp:74-134] for every trajectory, with
 * Segment::evaluate / Polynomial::evaluate [src/segment.cpp:51-58, polynomial.h:136-149].
 * The reference's serial recurrence (acc += dt, tau += dt, tau -= T_i on a strict
 * '>' crossing, loop counter restarting at the START of the segment holding
 * t_start) is replayed bit-exactly, so n_samples, sampling_times and segment_idx
 * equal the reference's; sample values use fused multiply-adds (value parity).
 *  t_start, t_end, dt [B]   per-trajectory range (the reference takes scalars per call)
 *  samples        [max_samples][D]  out or NULL   (rows >= n_samples[b]: not written with device
 *                                   pointers; UNSPECIFIED content with host pointers, where whole
 *                                   staging records are copied back; the same holds for the other
 *                                   per-sample outputs of this call and of mtg_feasibility_batch)
 *  sampling_times [max_samples]     out or NULL   (the reference's accumulated_time)
 *  segment_idx    [max_samples]     out or NULL
 *  n_samples      [B] int32         out or NULL
 * status: MTG_ST_OUT_OF_RANGE if t_start is beyond the trajectory or dt <= 0
 * (reference: LOG(ERROR) + empty result; t_start == max time is UB there, an error
 * here); MTG_ST_TRUNCATED if more than max_samples samples exist. */
int mtg_eval_range_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                         const double* seg_times, const double* t_start, const double* t_end,
                         const double* dt, int derivative, int max_samples, double* samples,
                         double* sampling_times, int32_t* segment_idx, int32_t* n_samples,
                         uint32_t* status, void* stream);

/* mtg_eval_at_batch: Trajectory::evaluate(t, derivative) [src/trajectory.cpp:41-72]
 * at M query times per trajectory. t [M], out [M][D], segment_idx [M] (or NULL;
 * -1 where t is out of range, in which case out is zero like the reference). */
int mtg_eval_at_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                      const double* seg_times, const double* t, int M, int derivative, double* out,
                      int32_t* segment_idx, uint32_t* status, void* stream);

/* mtg_feasibility_batch: fused sweep over the evaluateRange sample set: position,
 * velocity and acceleration of every sample, per-sample flags and per-trajectory
 * maxima. Sampled forms of: the v/a limit check [test_utils.h:43-54,
 * impl/polynomial_optimization_nonlinear_impl.h:2686-2733] and the tube / end-cap
 * geometry that the reference only hands to MOSEK
 * [impl/polynomial_optimization_qcqp_impl.h:369-474] (PARITY UNPINNED by any
 * reference test; D = 3 only).
 *  positions [K+1][3], radii [K][2] (pair<first,second> per segment, QC_H:55): tube
 *            check inputs, both NULL to skip it (bit2 is then always set)
 *  samples [max_samples][D] out or NULL (positions)
 *  flags   [max_samples] uint8 out or NULL: bit0 |v| <= v_max, bit1 |a| <= a_max, bit2 in tube
 *  max_v, max_a [B] out or NULL; feasible [B] uint8 out or NULL (all samples carry all bits) */
int mtg_feasibility_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                          const double* seg_times, const double* positions, const double* radii,
                          double v_max, double a_max, const double* t_start, const double* t_end,
                          const double* dt, int max_samples, double* samples, uint8_t* flags,
                          double* max_v, double* max_a, uint8_t* feasible, int32_t* n_samples,
                          uint32_t* status, void* stream);

/* --------------------------------- N1: the non-linear objective over the free endpoint derivatives
 * (device pointers only: these calls sit inside optimiser loops and only enqueue work on `stream`; the
 *  canonical constraint pattern of mtg_solve_batch; tensors as there, free_constraints = d_p [D][K-1][N/2-1])
 *
 * mtg_cost_derivative_batch: getCostAndGradientDerivative [NL_I:1537-1606]:
 *   J_d  [B]                      sum_dim [d_f; d_p]^T R [d_f; d_p]  (no 1/2)
 *   grad [D][K-1][N/2-1]          2 R_pf d_f + 2 R_pp d_p, in the layout of free_constraints
 *   diag [K-1][N/2-1]             the diagonal of 2 R_pp (same for every dimension; a Jacobi preconditioner)
 * Evaluated segment by segment from H(T) = T^(1-2d) S H1 S; the reference forms the dense R (getR) per call.
 *
 * mtg_soft_constraint_gradient_batch: getCostAndGradientSoftConstraints [NL_I:2365-2423] (central != 0) or
 * ...Simple [:2425-2490] (forward): J_sc = evaluateMaximumMagnitudeAsSoftConstraint [:2735-2766] of the given
 * trajectory and its finite-difference gradient with respect to every free derivative,
 *   grad[q] = (J_sc(d_p + inc e_q) - J_sc(d_p - inc e_q)) / (2 inc)   or   (J_sc(d_p + inc e_q) - J_sc) / inc,
 * where the reference re-runs setFreeConstraints + rpoly over ALL segments per perturbation; here only the two
 * segments next to the perturbed vertex are re-evaluated (2 D (K-1)(N/2-1) two-segment root problems per
 * trajectory in one launch of the extrema kernel). coeffs [K][D][N] are the coefficients OF d_p
 * (mtg_set_free_constraints_batch). AoS layout. constraints: HOST arrays derivatives[n], limits[n], n <= 4.
 *
 * mtg_nl_descent_batch: a projected-gradient driver standing in for NLOPT on
 * objectiveFunctionFreeConstraints[AndCollision] [NL_I:1024-1284] without the collision term:
 *   f = w_d J_d + w_sc J_sc,   trial = clamp(x_acc - step_b * (w_d grad_d + w_sc grad_sc) / diag, -bound, +bound)
 * (diag only when precondition != 0; bound[k] = |limit| of the constraint on derivative k,
 * setFreeEndpointDerivativeHardConstraints [NL_I:2858-2905]; the start point is projected onto the bounds
 * first). A trial point is accepted iff f did not
 * increase; a rejected trial halves that trajectory's step and restarts from the last accepted point, so
 * the returned point never has a larger f than the start. `iterations` trial steps; free_constraints is
 * updated in place (the last accepted point), coeffs [K][D][N] receives its coefficients, cost_history
 * [iterations + 2][3][B] (or NULL): J_d, J_sc and accepted (1 / 0) of trial point 0..iterations, then of the
 * returned point. Everything is enqueued on `stream` (capturable in a CUDA graph once the tables of
 * (N, derivative) are resident). AoS layout. */
int mtg_cost_derivative_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* positions,
                              const double* end_derivatives, const double* seg_times,
                              const double* free_constraints, double* J_d, double* grad, double* diag,
                              uint32_t* status, void* stream);
int mtg_soft_constraint_gradient_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                                       const double* seg_times, int n_constraints, const int32_t* derivatives,
                                       const double* limits, double weight, double maximum_cost,
                                       double increment, int central, double* J_sc, double* grad,
                                       uint32_t* status, void* stream);
int mtg_nl_descent_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* positions,
                         const double* end_derivatives, const double* seg_times, double* free_constraints,
                         int n_constraints, const int32_t* derivatives, const double* limits, double w_d,
                         double w_sc, double soft_weight, double maximum_cost, double increment, double step,
                         int precondition, int iterations, double* coeffs, double* cost_history,
                         uint32_t* status, void* stream);

/* --------------------------------- N4: collision potential against a dense distance grid
 * mtg_collision_cost_batch: getCostAndGradientCollision [NL_I:1608-1780] with
 * getCostAndGradientPotentialOctree [:1783-1917] and getCostPotential [:2659-2684] for every trajectory of the
 * batch in one sweep. The supereight octree of the reference (findOccupiedVoxels / getDistanceOctree,
 * :1920-2043: distance to the nearest occupied voxel inside a 20^3 window) is replaced by a DENSE grid of
 * distances that the caller precomputes once per map:
 *  grid [size0][size1][size2] double, DEVICE pointer shared by the batch: distance in metres at voxel
 *       (origin0 + i0, origin1 + i1, origin2 + i2); a voxel outside the grid counts as "no obstacle in reach"
 *       (the reference's DBL_MAX); the voxel of a position is (x / map_resolution) truncated toward zero
 *       like Eigen's cast<int>() [:1812]
 *  min_bound, max_bound [3]: positions within map_resolution of them are invalid = in collision [:1800-1807]
 *  J_c [B]: sum over the consulted samples of c(x) |v| time_sum; 0 when the trajectory collides [:1774]
 *  grad [3][K-1][N/2-1] (layout of free_constraints) or NULL: equation (14) [:1735-1748] with central
 *       differences of the potential over the six neighbour voxels. As in the reference a collision keeps the
 *       gradient accumulated up to the colliding sample (its zeroing loop iterates by value, :1776-1777).
 *  in_collision [B] uint8, n_checks [B] int32 (samples at which the map was consulted) out or NULL
 * Sampling per segment `for (t = 0; t < T_i; t += coll_check_time_increment)`, samples skipped until the
 * travelled distance reaches map_resolution [:1705-1708]. D = 3, canonical constraint pattern, device pointers. */
int mtg_collision_cost_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                             const double* seg_times, const double* grid, const int32_t grid_size[3],
                             const int32_t grid_origin_voxel[3], double map_resolution, const double min_bound[3],
                             const double max_bound[3], double coll_check_time_increment, double epsilon,
                             double robot_radius, double coll_pot_multiplier, double* J_c, double* grad,
                             uint8_t* in_collision, int32_t* n_checks, uint32_t* status, void* stream);

/* --------------------------------- N3: batched trajectory composition and I/O
 * mtg_vertex_at_time_batch: Trajectory::getVertexAtTime(t, max_derivative_order) [src/trajectory.cpp:248-254]
 * (getStartVertex / getGoalVertex with t = 0 / max time, :256-262): evaluate(t, k) for k = 0..max at one
 * time per trajectory. t [B]; out [max+1][D] records; segment_idx [B] or NULL (-1 and a zero vertex when t is
 * out of range, MTG_ST_OUT_OF_RANGE).
 *
 * mtg_pick_dimensions_batch: getTrajectoryWithSingleDimension / getTrajectoryWithAppendedDimension
 * [src/trajectory.cpp:136-182, src/segment.cpp:186-222] in one gather: output dimension q of every segment is
 * dimension pick[q] of trajectory set a (desc.D dimensions) when pick[q] < desc.D, else dimension
 * pick[q] - desc.D of set b (D_b dimensions; coeffs_b NULL when D_b = 0). pick: HOST array, n_out <= 8.
 * coeffs_out [K][n_out][N] records. Segment times are shared (the reference CHECKs equal K; equal times are
 * the caller's contract there too).
 *
 * mtg_concat_segments_batch: Trajectory::addTrajectories [src/trajectory.cpp:230-246] for whole batches: the
 * K_in[q] segments of trajectory b of every input set, one set after the other, into records of
 * sum K_in segments (coefficients and segment times). Strided DMA copies, no kernel.
 *
 * mtg_compute_cost_batch: PolynomialOptimization::computeCost [LIN_I:113-130] of GIVEN coefficients and
 * segment times (0.5 sum c^T Q(T) c, derivative desc.derivative_to_optimize) — what the reference returns
 * after updateSegmentTimes without a new solve.
 *
 * mtg_sample_dump_batch: the sample matrix of printMatlabSampledTrajectory [NL_I:2907-3003]: per segment
 * `for (t = 0; t < T_i; t += dt)` (the accumulation is replayed, so the row count is the reference's), rows
 * [t + segment start, pos(D), vel(D), acc(D), jerk(D), snap(D), tm]; rows [B][max_rows][5 D + 2], trajectory-
 * contiguous in both layouts; rows without a sample are zero and tm of ROW i holds the end time of segment i
 * like the reference. n_rows [B] or NULL; MTG_ST_TRUNCATED when max_rows is too small. */
int mtg_vertex_at_time_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                             const double* seg_times, const double* t, int max_derivative_order, double* out,
                             int32_t* segment_idx, uint32_t* status, void* stream);
int mtg_pick_dimensions_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs_a, int D_b,
                              const double* coeffs_b, int n_out, const int32_t* pick, double* coeffs_out,
                              void* stream);
int mtg_concat_segments_batch(mtg_ctx* ctx, int B, int D, int N, int memory, int layout, int n_inputs,
                              const int32_t* K_in, const double* const* coeffs_in, const double* const* times_in,
                              double* coeffs_out, double* times_out, void* stream);
int mtg_compute_cost_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                           const double* seg_times, double* cost, uint32_t* status, void* stream);
int mtg_sample_dump_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                          const double* seg_times, double dt, int max_rows, double* rows, int32_t* n_rows,
                          uint32_t* status, void* stream);

/* --------------------------------- N2: Bezier control points and the corridor constraints on them
 * mtg_control_points_batch: the control points of every segment — setupInverseControlPointMappingMatrix
 * + the extraction F B_inv C [d_f; d_p] of setupControlPointConstraints
 * [impl/polynomial_optimization_qcqp_impl.h:267-355] — and the VALUES of the reference's tube, end-cap and
 * sphere constraints on them [:357-474] (the reference only hands their coefficients to MOSEK; here a batch
 * of given trajectories is screened). A polynomial lies in the convex hull of its control points and the
 * tube-and-caps region is convex, so `feasible` is a sufficient (conservative) form of the sampled
 * predicate of mtg_feasibility_batch. Entries of B_inv in (-1e-5, 1e-5) are zeroed like the reference (:300-306).
 *  coeffs      [K][D][N]      in or NULL  (used when derivatives is NULL: endpoint derivatives are
 *                                          re-evaluated from the coefficients)
 *  derivatives [K+1][N/2][D]  in or NULL  C [d_f; d_p]: the full endpoint derivatives of every vertex
 *  positions [K+1][3], radii [K][2]  in (NULL when only control points are wanted): as mtg_feasibility_batch
 *  control_points [K][N][D]   out or NULL
 *  tube, cap_start, cap_end [K][N-2]  out or NULL (D = 3): |A x + b|^2 - r_tube^2, (-n).(x - p_start),
 *                              n.(x - p_end) on control points 1..N-2; feasible <=> value <= 0
 *  sphere [K]                 out or NULL: |x - v_{i+1}|^2 - radii[i].second^2 on the last control point;
 *                              -infinity for the last segment (not constrained by the reference, :349-351)
 *  max_value [B], feasible [B] uint8  out or NULL: the largest value / all values <= 0 */
int mtg_control_points_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                             const double* derivatives, const double* seg_times, const double* positions,
                             const double* radii, double* control_points, double* tube, double* cap_start,
                             double* cap_end, double* sphere, double* max_value, uint8_t* feasible,
                             uint32_t* status, void* stream);

/* --------------------------------- E6 / R1: analytic extrema of |p^(derivative)(t)|
 * Trajectory::computeMinMaxMagnitude(derivative, all dimensions, &minimum, &maximum)
 * [src/trajectory.cpp:184-220] for every trajectory, with
 * Segment::computeMinMaxMagnitudeCandidates / selectMinMaxMagnitudeFromCandidates
 * [src/segment.cpp:82-184], Polynomial::convolve / computeMinMaxCandidates
 * [src/polynomial.cpp:32-83, 163-181] and the Jenkins-Traub root finder
 * [src/rpoly/rpoly_ak1.cpp:57-937] behind it; also what
 * PolynomialOptimization::computeMaximumOfMagnitude [LIN_I:455-487] and the v/a limit
 * check evaluateMaximumMagnitudeConstraint [NL_I:2686-2733] consume (seg_max_*).
 * The candidate times of a segment are t = 0, t = T and the real roots in [0, T] of
 * g = sum_dim p_dim^(d) * p_dim^(d+1) (one dimension: of p^(d+1)); candidate order and tie
 * rules are the reference's (first candidate / earliest segment wins). Times are relative to
 * the segment start (extremum.h:41-42). The real roots are isolated by bounded Bernstein-basis
 * subdivision (variation-diminishing sign count) + bracketed Newton instead of a port of rpoly:
 * extremum VALUES agree with the reference to rounding; extremum TIMES to the root accuracy of
 * either method, except where a root is (numerically) multiple, e.g. at the rest-to-rest ends
 * (SURVEY.md section 7.4). Coefficients of g below 1e-12 of its scale on the segment count as zero.
 *  min_value, min_time, max_value, max_time [B] double; min_seg, max_seg [B] int32  out or NULL
 *  seg_max_value, seg_max_time [K]   out or NULL, maximum of every segment
 * status: MTG_ST_NO_CONVERGENCE if a bracketed refinement hit its iteration cap or the interval
 * stack of a warp was full (an interval with several roots was then taken as one bracket). */
int mtg_extrema_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                      const double* seg_times, int derivative, double* min_value, double* min_time,
                      int32_t* min_seg, double* max_value, double* max_time, int32_t* max_seg,
                      double* seg_max_value, double* seg_max_time, uint32_t* status, void* stream);

/* mtg_extrema_candidates_batch: the candidate LISTS behind the extrema — E6:
 * Segment::computeMinMaxMagnitudeCandidateTimes / computeMinMaxMagnitudeCandidates [src/segment.cpp:82-158],
 * i.e. what PolynomialOptimization::computeSegmentMaximumMagnitudeCandidates [LIN_I:396-417] returns,
 * and, with D = 1 (or a one-bit dim_mask), Polynomial::computeMinMaxCandidates /
 * selectMinMaxCandidatesFromRoots [src/polynomial.cpp:32-83] with the candidate values of
 * selectMinMaxFromCandidates [:116-143] — for every segment of every trajectory.
 *  t_start, t_end [K]   in or NULL: the interval per segment (NULL = 0 and the segment time)
 *  dim_mask             bit q set = dimension q takes part (`dimensions`, segment.cpp:82-86); 0 = all
 *  cand_time  [K][max_candidates]  out or NULL: t_start, t_end, then the real sign-changing roots of
 *             g in [t_start, t_end] in ASCENDING order (the reference lists them in Jenkins-Traub's
 *             order of discovery, and also keeps even-multiplicity roots, which cannot be extrema)
 *  cand_value [K][max_candidates]  out or NULL: |p^(derivative)(t)| over the selected dimensions
 *  n_candidates [K] int32          out or NULL: candidates of the segment (entries beyond
 *             max_candidates are dropped and MTG_ST_TRUNCATED is set)
 *
 * mtg_poly_real_roots_batch: R1 root-list interface — the real roots inside [t_lo, t_hi] of B
 * polynomials with n_coeffs coefficients each (increasing powers): findRootsJenkinsTraub
 * [src/rpoly/rpoly_ak1.cpp:70-117] followed by the real / in-range selection of
 * selectMinMaxCandidatesFromRoots [src/polynomial.cpp:46-60]. Leading zero coefficients are
 * stripped like findLastNonZeroCoeff [:57-68]; a constant polynomial has no roots. Roots come
 * out ascending; only sign-changing roots are found (see mtg_extrema_batch). coeffs [n_coeffs]
 * records in `layout`, t_lo / t_hi [B], roots [max_roots] records, n_roots [B].
 *
 * mtg_soft_constraint_batch: E5 soft form — evaluateMaximumMagnitudeAsSoftConstraint
 * [NL_I:2735-2766] on top of evaluateMaximumMagnitudeConstraint [NL_I:2686-2733]: for the
 * constraints c = (derivatives[c], limits[c]) (HOST arrays)
 *   violations[c][b] = max_t |p^(derivatives[c])(t)| - limits[c]          (out or NULL, [n][B])
 *   cost[b] = sum_c min(maximum_cost, exp(violations[c][b] / limits[c] * weight))
 * Device pointers only (it sits inside optimiser loops). */
int mtg_extrema_candidates_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                                 const double* seg_times, const double* t_start, const double* t_end,
                                 int derivative, int dim_mask, int max_candidates, double* cand_time,
                                 double* cand_value, int32_t* n_candidates, uint32_t* status, void* stream);
int mtg_poly_real_roots_batch(mtg_ctx* ctx, int B, int n_coeffs, int memory, int layout, const double* coeffs,
                              const double* t_lo, const double* t_hi, int max_roots, double* roots,
                              int32_t* n_roots, uint32_t* status, void* stream);
int mtg_soft_constraint_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                              const double* seg_times, int n_constraints, const int32_t* derivatives,
                              const double* limits, double weight, double maximum_cost, double* cost,
                              double* violations, uint32_t* status, void* stream);

/* ------------------------------------------- candidate generator of a sweep (G1 on the device)
 * mtg_generate_candidates_batch: createRandomVertices [src/vertex.cpp:27-82] (uniform positions in
 * [pos_min, pos_max] per dimension, resampled until |pos - last| > 0.2, :65-72) + estimateSegmentTimesNfabian
 * [:252-269] for B candidates, written straight into device tensors positions [K+1][D] / seg_times [K].
 * The reference's std::mt19937 is serial; a sharded sweep needs a counter-based generator: Philox4x32-10, key =
 * seed, counter = (candidate index first_index + b, draw number, block) — candidate b is identical on every
 * rank and for every batch split; draw n yields the D uniforms of one attempt (53-bit, u0 = words 0,1 of block 0,
 * u1 = words 2,3, u2 / u3 from block 1), position = pos_min + u * (pos_max - pos_min) (multiply, then add).
 * pos_min / pos_max: HOST arrays [D]. magic_fabian_constant: 6.5 in the reference (vertex.h:128). */
int mtg_generate_candidates_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, uint64_t seed, int64_t first_index,
                                  const double* pos_min, const double* pos_max, double v_max, double a_max,
                                  double magic_fabian_constant, double* positions, double* seg_times, void* stream);

/* ------------------------------------------- candidate sweep: argmin of computeCost()
 * The reference picks the best of many candidate trajectories on the host, one
 * computeCost() [LIN_I:113-130] at a time (e.g. the random restarts of
 * PolynomialOptimizationNonLinear, NL_I:274-330). Sharded over GPUs, the sweep needs
 * exactly one exchange: every rank's {cost, global index} pair (16 bytes).
 *
 * mtg_argmin_batch: device-side argmin over cost[0..n) (DEVICE pointers). Entries whose
 * status is non-zero (status may be NULL) or whose cost is NaN never win; ties go to the
 * lower global index = global_offset + i. best -> device struct {double cost; int64 idx}
 * (idx = -1 when nothing qualified); accumulate != 0 folds the pair already stored in
 * *best into the result (running argmin over several batches).
 *
 * mtg_nccl_unique_id / mtg_nccl_init: one NCCL communicator per context (ncclGetUniqueId
 * on rank 0, the 128 id bytes handed to every rank by the host's own channel, then
 * ncclCommInitRank). NCCL is dlopen()ed on first use (MTG_NCCL_LIB, else libnccl.so.2).
 *
 * mtg_argmin_allgather: local argmin + ncclAllGather of the pairs over NVLink/NVSwitch +
 * final selection; returns the same (cost, index) on every rank through HOST pointers
 * (the stream is synchronised). Without an initialised communicator it is the local argmin.
 *
 * mtg_best_allgather: the same exchange WITHOUT a host round trip, for sweeps that keep a running
 * best on the device (mtg_argmin_batch with accumulate): all-gathers the DEVICE pair best_local of
 * every rank (ncclAllGather, 16 bytes per rank, enqueued on `stream`) and folds the pairs on the
 * device into the DEVICE pair best_global — identical on every rank; nothing is synchronised, the
 * caller reads best_global when it needs it. Without an initialised communicator: a copy. */
int mtg_argmin_batch(mtg_ctx* ctx, const double* cost, const uint32_t* status, int64_t n,
                     int64_t global_offset, int accumulate, void* best, void* stream);
/* mtg_set_solve_overlap: lets consecutive DEVICE-memory solves of one stream overlap (programmatic dependent
 * launch, sm_90+). 65,536 solves are 3.46 waves of CTAs: with the option on, mtg_solve_batch / mtg_solve_argmin_batch
 * are launched with programmatic stream serialization, so the next solve's CTAs start on the SMs the last, partial
 * wave of the previous one leaves empty; they run the part that only READS global memory (the forward elimination)
 * and wait for the previous kernel to finish before their first global store. Outputs may therefore be reused from
 * call to call as usual. CONTRACT while it is on: the inputs of a solve (positions, end_derivatives, seg_times) must
 * not be written by the stream operation immediately preceding it (e.g. do not put mtg_generate_candidates_batch
 * directly before the solve of its own output; generate one batch ahead instead). Off by default. */
int mtg_set_solve_overlap(mtg_ctx* ctx, int enabled);

/* mtg_solve_argmin_batch: one step of a candidate sweep in ONE launch — mtg_solve_batch (same
 * arguments and meaning; coeffs, cost, free_constraints and status may ALL be NULL) with
 * mtg_argmin_batch over its costs folded into the solve kernel's epilogue: every CTA publishes the
 * best {cost, global index} of its trajectories and the last CTA folds them, plus the pair already
 * in *best when accumulate != 0, into the DEVICE pair *best. Same total order and exclusions as
 * mtg_argmin_batch (failed solves and NaN costs never win, ties -> lower global index =
 * global_offset + b). This is the loop body of the reference's restart loop — solveLinear() +
 * computeCost() + keep the cheapest [NL_I:274-330, LIN_I:113-130, 337-379] — without the cost vector
 * ever leaving the SMs. Device pointers only. Shapes the fused epilogue does not cover (K = 1,
 * chains so long that a CTA is not whole warps) run as the two launches it replaces. */
int mtg_solve_argmin_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* positions,
                           const double* end_derivatives, const double* seg_times, double* coeffs, double* cost,
                           double* free_constraints, uint32_t* status, int64_t global_offset, int accumulate,
                           void* best, void* stream);
int mtg_nccl_unique_id(mtg_ctx* ctx, uint8_t id[128]);
int mtg_nccl_init(mtg_ctx* ctx, const uint8_t id[128], int rank, int world);
int mtg_argmin_allgather(mtg_ctx* ctx, const double* cost, const uint32_t* status, int64_t n_local,
                         int64_t global_offset, double* best_cost, int64_t* best_idx, void* stream);
int mtg_best_allgather(mtg_ctx* ctx, const void* best_local, void* best_global, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MTG_CUDA_H_ */

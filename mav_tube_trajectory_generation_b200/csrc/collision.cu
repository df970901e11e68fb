// libmtg_cuda.so — N4: the collision potential of the non-linear layer against a DENSE distance grid, fused
// into one sample sweep per trajectory (the reference walks a supereight octree per sample).
//
// Replaces (reference, impl/polynomial_optimization_nonlinear_impl.h = NL_I):
//   getCostAndGradientCollision        NL_I:1608-1780   line integral of the potential along the trajectory:
//       per segment `for (t = 0; t < T_i; t += dt)`; position and velocity; samples are skipped until the
//       travelled distance reaches map_resolution; then  J_c += c(x) |v| time_sum  and (equation (14))
//       grad_c[k] += |v| time_sum dc/dx_k (T^T L_pp) + time_sum c v_k / |v| (T^T V L_pp);
//       a sample in collision ends the sweep with J_c = 0
//   getCostAndGradientPotentialOctree  NL_I:1783-1917   distance at the sample's voxel and at its six neighbours
//       (central differences of the potential, / (2 map_resolution)); outside [min_bound + res, max_bound - res]
//       the distance is 0 (= collision)
//   getCostPotential                   NL_I:2659-2684   the potential of a distance
//   findOccupiedVoxels / getDistanceOctree NL_I:1920-2043: replaced by a lookup in a dense grid of distances
//       (metres, one double per voxel) that the caller precomputes — the map format decision of SURVEY (f) N4.
//
// T^T L_pp never needs the 100 x 36 matrix L = A^-1 M: for a sample of segment i only the free derivatives of
// vertices i and i + 1 have non-zero columns, and (T^T A_i^-1)[col] = T_i^k P_col(t / T_i) with the constant
// polynomials P_col(u) = sum_j A(1)^-1[j][col] u^j (the velocity row is T_i^(k-1) P_col'(u)). One thread per
// trajectory replays the reference's serial state machine (time_sum, dist_sum, prev_pos); the gradient of a
// segment accumulates in registers and is flushed once per segment.
#include "host_common.h"
#include "solve_canonical.cuh"  // at<AOS>()

MTG_REGISTER_TABLES()

using namespace mtg;

namespace {

struct CollisionParams {
  const double* __restrict__ coeffs;     // elem ((i*D + dim)*N + j), rec K*3*N
  const double* __restrict__ seg_times;  // elem i, rec K
  const double* __restrict__ grid;       // [nx][ny][nz] distances in metres (shared by the batch)
  double* __restrict__ J_c;              // [B]
  double* __restrict__ grad;             // elem ((dim*(K-1) + v-1)*NF + k-1), rec 3*(K-1)*NF; or nullptr
  uint8_t* __restrict__ in_collision;    // [B] or nullptr
  int32_t* __restrict__ n_checks;        // [B] or nullptr: samples at which the map was consulted
  uint32_t* __restrict__ status;         // [B] or nullptr
  int nx, ny, nz, ox, oy, oz;            // grid size and the voxel index of grid[0][0][0]
  double res, dt;                        // map_resolution, coll_check_time_increment
  double min_bound[3], max_bound[3];
  double epsilon, robot_radius, multiplier;
  int B, b0, nb, K, N;
};

// getCostPotential, NL_I:2659-2684
__device__ __forceinline__ double cost_potential(const CollisionParams& p, double distance, bool* collision) {
  *collision = false;
  distance -= p.robot_radius;
  if (distance <= 0.0) {
    *collision = true;
    return p.multiplier * (-distance) + 0.5 * p.epsilon;
  }
  if (distance <= p.epsilon) {
    const double e = distance - p.epsilon;
    return 0.5 * 1.0 / p.epsilon * e * e;
  }
  return 0.0;
}

// distance of voxel (vx, vy, vz); outside the grid: "no occupied voxel in reach" (the reference's DBL_MAX * res)
__device__ __forceinline__ double grid_distance(const CollisionParams& p, int vx, int vy, int vz) {
  const int ix = vx - p.ox, iy = vy - p.oy, iz = vz - p.oz;
  if (ix < 0 || iy < 0 || iz < 0 || ix >= p.nx || iy >= p.ny || iz >= p.nz) return 1.7976931348623157e308;
  return __ldg(p.grid + ((size_t)ix * p.ny + iy) * p.nz + iz);
}

template <int HN, bool AOS>
__global__ void __launch_bounds__(128) collision_kernel(const CollisionParams p) {
  constexpr int N = 2 * HN, NF = HN - 1, D = 3;
  const int local = blockIdx.x * blockDim.x + threadIdx.x;
  if (local >= p.nb) return;
  const int b = p.b0 + local;
  const size_t B = (size_t)p.B;
  const int K = p.K;
  const size_t rec_c = (size_t)K * D * N, rec_g = (size_t)D * (K - 1) * NF;
  if (p.grad)
    for (int e = 0; e < D * (K - 1) * NF; ++e) p.grad[at<AOS>((size_t)e, rec_g, B, b)] = 0.0;
  double J = 0.0;
  bool collided = false;
  uint32_t st = 0;
  int checks = 0;
  // numerical integral state (NL_I:1656-1661)
  double prev[3] = {0.0, 0.0, 0.0};
  double time_sum = -1.0, dist_sum = 0.0;
  for (int i = 0; i < K && !collided; ++i) {
    double T = p.seg_times[at<AOS>((size_t)i, (size_t)K, B, b)];
    if (!(T > 0.0) || !(T < 1.7e308)) {
      st |= 1u;
      T = 1.0;
    }
    double c[D][N];
#pragma unroll
    for (int dim = 0; dim < D; ++dim)
#pragma unroll
      for (int j = 0; j < N; ++j) c[dim][j] = p.coeffs[at<AOS>((size_t)(i * D + dim) * N + j, rec_c, B, b)];
    double gs[D][NF], ge[D][NF];  // gradient of this segment w.r.t. the free derivatives of vertex i / i + 1
#pragma unroll
    for (int dim = 0; dim < D; ++dim)
#pragma unroll
      for (int k = 0; k < NF; ++k) gs[dim][k] = ge[dim][k] = 0.0;
    double tp[HN];  // T^k
    tp[0] = 1.0;
#pragma unroll
    for (int k = 1; k < HN; ++k) tp[k] = tp[k - 1] * T;
    const double invT = 1.0 / T;
    double t = 0.0;
    for (t = 0.0; t < T; t += p.dt) {
      double pos[3], vel[3];
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
        double x = c[dim][N - 1], v = 0.0;
#pragma unroll
        for (int j = N - 2; j >= 0; --j) {
          v = fma(v, t, x);
          x = fma(x, t, c[dim][j]);
        }
        pos[dim] = x;
        vel[dim] = v;
      }
      if (time_sum < 0.0) {  // the very first sample only primes the integral
        time_sum = 0.0;
        prev[0] = pos[0]; prev[1] = pos[1]; prev[2] = pos[2];
        continue;
      }
      time_sum += p.dt;
      {
        const double dx = pos[0] - prev[0], dy = pos[1] - prev[1], dz = pos[2] - prev[2];
        dist_sum += sqrt(dx * dx + dy * dy + dz * dz);
      }
      prev[0] = pos[0]; prev[1] = pos[1]; prev[2] = pos[2];
      if (dist_sum < p.res) continue;
      ++checks;
      // getCostAndGradientPotentialOctree
      const bool valid = !(pos[0] < p.min_bound[0] + p.res || pos[0] > p.max_bound[0] - p.res ||
                           pos[1] < p.min_bound[1] + p.res || pos[1] > p.max_bound[1] - p.res ||
                           pos[2] < p.min_bound[2] + p.res || pos[2] > p.max_bound[2] - p.res);
      const int vx = (int)(pos[0] / p.res), vy = (int)(pos[1] / p.res), vz = (int)(pos[2] / p.res);  // cast<int>: toward zero
      bool hit = false;
      const double cost = cost_potential(p, valid ? grid_distance(p, vx, vy, vz) : 0.0, &hit);
      if (hit) {
        collided = true;
        break;
      }
      const double nv = sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
      J += cost * nv * time_sum;
      if (p.grad && nv > 1e-6) {
        double gc[3];
        bool dummy;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double dl = 0.0, dr = 0.0;  // left_dist / right_dist stay unset in the reference when !valid; valid here (no break above)
          if (valid) {
            dl = grid_distance(p, vx - (k == 0), vy - (k == 1), vz - (k == 2));
            dr = grid_distance(p, vx + (k == 0), vy + (k == 1), vz + (k == 2));
          }
          gc[k] = (cost_potential(p, dr, &dummy) - cost_potential(p, dl, &dummy)) / (2.0 * p.res);
        }
        const double u = t * invT;
#pragma unroll
        for (int k = 1; k < HN; ++k) {
          // P_col(u) and P_col'(u) for the start column k and the end column h + k
          double ps = MTG_AI(N - 1, k), dps = 0.0, pe = MTG_AI(N - 1, HN + k), dpe = 0.0;
#pragma unroll
          for (int j = N - 2; j >= 0; --j) {
            dps = fma(dps, u, ps);
            ps = fma(ps, u, MTG_AI(j, k));
            dpe = fma(dpe, u, pe);
            pe = fma(pe, u, MTG_AI(j, HN + k));
          }
          const double ws = tp[k] * ps, we = tp[k] * pe;                  // (T^T A^-1)[col]
          const double vs = tp[k - 1] * dps, ve = tp[k - 1] * dpe;        // (T^T V A^-1)[col]
#pragma unroll
          for (int dim = 0; dim < D; ++dim) {
            const double a1 = nv * time_sum * gc[dim], a2 = time_sum * cost * vel[dim] / nv;
            gs[dim][k - 1] += a1 * ws + a2 * vs;
            ge[dim][k - 1] += a1 * we + a2 * ve;
          }
        }
      }
      dist_sum = 0.0;
      time_sum = 0.0;
      prev[0] = pos[0]; prev[1] = pos[1]; prev[2] = pos[2];
    }
    if (p.grad) {  // flush this segment's share (also after a collision: the reference keeps what it accumulated)
#pragma unroll
      for (int dim = 0; dim < D; ++dim)
#pragma unroll
        for (int k = 0; k < NF; ++k) {
          if (i >= 1) p.grad[at<AOS>((size_t)(dim * (K - 1) + (i - 1)) * NF + k, rec_g, B, b)] += gs[dim][k];
          if (i + 1 <= K - 1) p.grad[at<AOS>((size_t)(dim * (K - 1) + i) * NF + k, rec_g, B, b)] += ge[dim][k];
        }
    }
    if (collided) break;
    time_sum += -p.dt + (T - t);  // NL_I:1757
  }
  p.J_c[b] = collided ? 0.0 : J;  // NL_I:1774-1778
  if (p.in_collision) p.in_collision[b] = collided ? 1 : 0;
  if (p.n_checks) p.n_checks[b] = checks;
  if (p.status) p.status[b] = st;
}

}  // namespace

extern "C" int mtg_collision_cost_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                                        const double* seg_times, const double* grid, const int32_t grid_size[3],
                                        const int32_t grid_origin_voxel[3], double map_resolution,
                                        const double min_bound[3], const double max_bound[3],
                                        double coll_check_time_increment, double epsilon, double robot_radius,
                                        double coll_pot_multiplier, double* J_c, double* grad, uint8_t* in_collision,
                                        int32_t* n_checks, uint32_t* status, void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if (desc->memory != MTG_MEM_DEVICE)
    return fail(ctx, MTG_ERR_UNSUPPORTED, "mtg_collision_cost_batch takes device pointers (it runs inside optimiser loops)");
  if (desc->D != 3) return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "the collision term is 3-D (NL_I:1800-1807)");
  if (desc->N < 4) return fail(ctx, MTG_ERR_UNSUPPORTED, "supported N: {4,6,8,10,12}");
  if (!coeffs || !seg_times || !grid || !grid_size || !grid_origin_voxel || !min_bound || !max_bound || !J_c ||
      !(map_resolution > 0.0) || !(coll_check_time_increment > 0.0) || !(epsilon > 0.0) || grid_size[0] < 1 ||
      grid_size[1] < 1 || grid_size[2] < 1)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "coeffs, seg_times, a non-empty grid, bounds, J_c, map_resolution > 0, "
                                               "coll_check_time_increment > 0 and epsilon > 0 are required");
  if (grad && desc->K < 2) grad = nullptr;  // a single segment has no free derivatives
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream_;
  TableGuard tables(ctx, desc->N, desc->derivative_to_optimize, s);  // A(1)^-1 (independent of the cost derivative)
  if (tables.rc()) return tables.rc();
  CollisionParams p = {};
  p.coeffs = coeffs; p.seg_times = seg_times; p.grid = grid; p.J_c = J_c; p.grad = grad;
  p.in_collision = in_collision; p.n_checks = n_checks; p.status = status;
  p.nx = grid_size[0]; p.ny = grid_size[1]; p.nz = grid_size[2];
  p.ox = grid_origin_voxel[0]; p.oy = grid_origin_voxel[1]; p.oz = grid_origin_voxel[2];
  p.res = map_resolution; p.dt = coll_check_time_increment;
  for (int k = 0; k < 3; ++k) {
    p.min_bound[k] = min_bound[k];
    p.max_bound[k] = max_bound[k];
  }
  p.epsilon = epsilon; p.robot_radius = robot_radius; p.multiplier = coll_pot_multiplier;
  p.B = desc->B; p.b0 = 0; p.nb = desc->B; p.K = desc->K; p.N = desc->N;
  const int grid_dim = (p.nb + 127) / 128;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
#define MTG_LAUNCH_COLL(HN_)                                                       \
  if (aos) collision_kernel<HN_, true><<<grid_dim, 128, 0, s>>>(p);                 \
  else collision_kernel<HN_, false><<<grid_dim, 128, 0, s>>>(p)
  switch (desc->N) {
    case 4: MTG_LAUNCH_COLL(2); break;
    case 6: MTG_LAUNCH_COLL(3); break;
    case 8: MTG_LAUNCH_COLL(4); break;
    case 10: MTG_LAUNCH_COLL(5); break;
    case 12: MTG_LAUNCH_COLL(6); break;
    default: return fail(ctx, MTG_ERR_UNSUPPORTED, "supported N: {4,6,8,10,12}");
  }
#undef MTG_LAUNCH_COLL
  ++ctx->launches;
  MTG_CUDA_TRY(cudaGetLastError());
  return MTG_OK;
}

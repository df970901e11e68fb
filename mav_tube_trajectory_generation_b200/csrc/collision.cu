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
// polynomials P_col(u) = sum_j A(1)^-1[j][col] u^j (the velocity row is T_i^(k-1) P_col'(u)). One WARP per
// trajectory, lane = sample (see collision_kernel): the reference's serial state machine (time_sum, dist_sum,
// prev_pos) is replayed as written, everything else runs lane-parallel.
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

// Shared memory of one warp (doubles): segment coefficients [3][N] | powers of u per sample [32][N | 1] (after a
// segment: its moments [2][3][N]) | gradient weights per sample [32][7] | gradient [3][K-1][NF]
__host__ __device__ inline int coll_ldu(int N) { return N | 1; }
__host__ __device__ inline int coll_warp_doubles(int N, int K) {
  return 3 * N + 32 * coll_ldu(N) + 32 * 7 + 3 * (K - 1) * (N / 2 - 1);
}
constexpr int kCollWarps = 4;

// One WARP per trajectory, lane = sample, in chunks of 32 samples of one segment:
//  1. the lane's t by the reference's repeated `t += dt` (lane many adds from the chunk's first t: same bits);
//  2. position and velocity (coupled Horner, coefficients broadcast from shared memory);
//  3. the step |pos_k - pos_{k-1}| (neighbour by shuffle, the chunk's first from the carried previous sample);
//  4. the reference's gate — skip samples until the travelled distance reaches map_resolution — is a serial
//     recurrence over (time_sum, dist_sum): replayed by every lane over the 32 steps (uniform, 2 adds and a compare
//     per sample); a lane keeps time_sum if its own sample is a check;
//  5. the checks (grid lookups, potential, central differences) run lane-parallel; the first colliding check ends
//     the sweep, the checks before it still count for the gradient like in the reference;
//  6. equation (14): the reference evaluates the 2 NF column polynomials P_col(u), P_col'(u) per check. They are
//     linear in the powers of u, so the checks only accumulate the moments M1[dim][j] = sum a1_dim u^j,
//     M2[dim][j] = sum a2_dim u^j (lane = (dim, j), samples in order), and one small product per segment turns
//     them into the gradient: T^k sum_j A(1)^-1[j][col] M1[dim][j] + T^(k-1) sum_j j A(1)^-1[j][col] M2[dim][j-1].
template <int HN, bool AOS>
__global__ void __launch_bounds__(kCollWarps * 32) collision_kernel(const CollisionParams p) {
  constexpr int N = 2 * HN, NF = HN - 1, D = 3, NO = (D * N + 31) / 32;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ double coll_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int local = blockIdx.x * kCollWarps + warp;
  if (local >= p.nb) return;  // whole warp; no CTA-wide barrier below
  const int b = p.b0 + local;
  const size_t B = (size_t)p.B;
  const int K = p.K;
  const size_t rec_c = (size_t)K * D * N, rec_g = (size_t)D * (K - 1) * NF;
  constexpr int LDU = N | 1;
  double* s_c = coll_smem + (size_t)warp * coll_warp_doubles(N, K);
  double* s_up = s_c + D * N;
  double* s_a = s_up + 32 * LDU;
  double* s_grad = s_a + 32 * 7;
  const bool want_grad = p.grad != nullptr;
  const int n_grad = D * (K - 1) * NF;
  for (int e = lane; e < n_grad; e += 32) s_grad[e] = 0.0;

  double Jl = 0.0;  // this lane's share of J_c
  bool collided = false;
  uint32_t st = 0;
  int checks = 0;
  // numerical integral state (NL_I:1656-1661): the same values in every lane
  double prev[3] = {0.0, 0.0, 0.0};
  double time_sum = -1.0, dist_sum = 0.0;
  for (int i = 0; i < K && !collided; ++i) {
    double T = p.seg_times[at<AOS>((size_t)i, (size_t)K, B, b)];
    if (!(T > 0.0) || !(T < 1.7e308)) {
      st |= 1u;
      T = 1.0;
    }
    __syncwarp();
    for (int e = lane; e < D * N; e += 32) s_c[e] = p.coeffs[at<AOS>((size_t)i * D * N + e, rec_c, B, b)];
    __syncwarp();
    double tp[HN];  // T^k
    tp[0] = 1.0;
#pragma unroll
    for (int k = 1; k < HN; ++k) tp[k] = tp[k - 1] * T;
    const double invT = 1.0 / T;
    double M1[NO], M2[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) M1[o] = M2[o] = 0.0;
    double t0 = 0.0, t_exit = 0.0;
    for (bool seg_done = false; !seg_done && !collided;) {
      // 1. sample times
      double t = t0;
      for (int k = 0; k < 31; ++k)
        if (k < lane) t += p.dt;
      const int nval = __popc(__ballot_sync(FULL, t < T));  // the valid lanes are a prefix (dt > 0)
      if (nval < 32) {
        seg_done = true;
        t_exit = __shfl_sync(FULL, t, nval);  // the value the reference's loop variable ends with
      } else {
        t0 = __shfl_sync(FULL, t, 31) + p.dt;
      }
      if (nval == 0) break;
      // 2. position and velocity
      double pos[3], vel[3];
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
        double x = s_c[dim * N + N - 1], v = 0.0;
#pragma unroll
        for (int j = N - 2; j >= 0; --j) {
          v = fma(v, t, x);
          x = fma(x, t, s_c[dim * N + j]);
        }
        pos[dim] = x;
        vel[dim] = v;
      }
      // 3. distance to the previous sample
      double step;
      {
        double q0 = __shfl_up_sync(FULL, pos[0], 1), q1 = __shfl_up_sync(FULL, pos[1], 1),
               q2 = __shfl_up_sync(FULL, pos[2], 1);
        if (lane == 0) {
          q0 = prev[0];
          q1 = prev[1];
          q2 = prev[2];
        }
        const double dx = pos[0] - q0, dy = pos[1] - q1, dz = pos[2] - q2;
        step = sqrt(dx * dx + dy * dy + dz * dz);
      }
      prev[0] = __shfl_sync(FULL, pos[0], nval - 1);
      prev[1] = __shfl_sync(FULL, pos[1], nval - 1);
      prev[2] = __shfl_sync(FULL, pos[2], nval - 1);
      // 4. the gate (NL_I:1680-1700): which samples consult the map, and with which time_sum. Uniform: every lane
      // replays the same recurrence over the chunk's steps (broadcast from shared memory).
      s_a[lane * 7 + 6] = step;
      __syncwarp();
      unsigned gate = 0u;
      double my_ts = 0.0;
      {
        int k = 0;
        if (time_sum < 0.0) {  // the very first sample only primes the integral
          time_sum = 0.0;
          k = 1;
        }
#pragma unroll 4
        for (; k < nval; ++k) {
          time_sum += p.dt;
          dist_sum += s_a[k * 7 + 6];
          if (!(dist_sum < p.res)) {
            gate |= 1u << k;
            s_a[k * 7 + 6] = time_sum;  // the consumed step makes room for the check's time_sum (same value from every lane)
            dist_sum = 0.0;
            time_sum = 0.0;
          }
        }
      }
      __syncwarp();
      bool my_check = (gate >> lane) & 1u;
      if (my_check) my_ts = s_a[lane * 7 + 6];
      __syncwarp();  // the slot is rewritten by the next chunk
      // 5. the checks
      unsigned cm = gate;
      if (cm == 0) continue;
      bool valid = false, hit = false;
      int vx = 0, vy = 0, vz = 0;
      double cost = 0.0;
      if (my_check) {
        // getCostAndGradientPotentialOctree
        valid = !(pos[0] < p.min_bound[0] + p.res || pos[0] > p.max_bound[0] - p.res ||
                  pos[1] < p.min_bound[1] + p.res || pos[1] > p.max_bound[1] - p.res ||
                  pos[2] < p.min_bound[2] + p.res || pos[2] > p.max_bound[2] - p.res);
        vx = (int)(pos[0] / p.res);  // cast<int>: toward zero
        vy = (int)(pos[1] / p.res);
        vz = (int)(pos[2] / p.res);
        cost = cost_potential(p, valid ? grid_distance(p, vx, vy, vz) : 0.0, &hit);
      }
      const unsigned hm = __ballot_sync(FULL, my_check && hit);
      if (hm) {  // the first sample in collision ends the sweep; it is counted, the ones after it never happen
        const int first = __ffs(hm) - 1;
        collided = true;
        cm &= (2u << first) - 1u;
        if (lane >= first) my_check = false;
      }
      checks += __popc(cm);
      double a1[3] = {0.0, 0.0, 0.0}, a2[3] = {0.0, 0.0, 0.0};
      if (my_check) {
        const double nv = sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
        Jl += cost * nv * my_ts;
        if (want_grad && nv > 1e-6) {
          bool dummy;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            // (left_dist / right_dist stay unset in the reference when !valid; valid here: no collision above)
            double dl = 0.0, dr = 0.0;
            if (valid) {
              dl = grid_distance(p, vx - (k == 0), vy - (k == 1), vz - (k == 2));
              dr = grid_distance(p, vx + (k == 0), vy + (k == 1), vz + (k == 2));
            }
            const double gc = (cost_potential(p, dr, &dummy) - cost_potential(p, dl, &dummy)) / (2.0 * p.res);
            a1[k] = nv * my_ts * gc;
            a2[k] = my_ts * cost * vel[k] / nv;
          }
        }
      }
      if (want_grad) {
        // 6. moments of this chunk
        {
          const double u = t * invT;
          double up = 1.0;
#pragma unroll
          for (int j = 0; j < N; ++j) {
            s_up[lane * LDU + j] = up;
            up *= u;
          }
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            s_a[lane * 7 + k] = a1[k];
            s_a[lane * 7 + 3 + k] = a2[k];
          }
        }
        __syncwarp();
#pragma unroll
        for (int o = 0; o < NO; ++o) {
          const int e = lane + 32 * o;
          if (e < D * N) {
            const int dim = e / N, j = e - dim * N;
            double m1 = M1[o], m2 = M2[o];
            for (int smp = 0; smp < nval; ++smp) {
              const double w = s_up[smp * LDU + j];
              m1 = fma(s_a[smp * 7 + dim], w, m1);
              m2 = fma(s_a[smp * 7 + 3 + dim], w, m2);
            }
            M1[o] = m1;
            M2[o] = m2;
          }
        }
        __syncwarp();
      }
    }
    if (want_grad) {
      // this segment's share of the gradient (also after a collision: the reference keeps what it accumulated)
      double* s_m = s_up;  // [2][D][N]
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        const int e = lane + 32 * o;
        if (e < D * N) {
          s_m[e] = M1[o];
          s_m[D * N + e] = M2[o];
        }
      }
      __syncwarp();
      if (lane < D * 2 * NF) {
        const int dim = lane / (2 * NF), cc = lane - dim * 2 * NF;
        const bool end = cc >= NF;
        const int k = (end ? cc - NF : cc) + 1;    // derivative order of the free variable
        const int col = end ? HN + k : k;          // column of A(1)^-1
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const double aij = MTG_AI(j, col);
          s1 = fma(aij, s_m[dim * N + j], s1);
          if (j >= 1) s2 = fma((double)j * aij, s_m[D * N + dim * N + j - 1], s2);
        }
        double tk = 1.0, tk1 = 1.0;  // T^k, T^(k-1)
#pragma unroll
        for (int q = 1; q < HN; ++q) {
          if (q == k) tk = tp[q];
          if (q == k - 1) tk1 = tp[q];
        }
        const double val = tk * s1 + tk1 * s2;
        const int vtx = end ? i + 1 : i;  // the vertex whose free derivative this is
        if (vtx >= 1 && vtx <= K - 1) s_grad[(dim * (K - 1) + (vtx - 1)) * NF + k - 1] += val;
      }
      __syncwarp();
    }
    if (collided) break;
    time_sum += -p.dt + (T - t_exit);  // NL_I:1757
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) Jl += __shfl_xor_sync(FULL, Jl, m);
  if (lane == 0) {
    p.J_c[b] = collided ? 0.0 : Jl;  // NL_I:1774-1778
    if (p.in_collision) p.in_collision[b] = collided ? 1 : 0;
    if (p.n_checks) p.n_checks[b] = checks;
    if (p.status) p.status[b] = st;
  }
  if (want_grad) {
    __syncwarp();
    for (int e = lane; e < n_grad; e += 32) p.grad[at<AOS>((size_t)e, rec_g, B, b)] = s_grad[e];
  }
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
  const int grid_dim = (p.nb + kCollWarps - 1) / kCollWarps;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  const size_t smem = (size_t)kCollWarps * coll_warp_doubles(desc->N, desc->K) * sizeof(double);
  if (smem > ctx->smem_optin)
    return fail(ctx, MTG_ERR_UNSUPPORTED, "mtg_collision_cost_batch: K too large for the per-warp gradient in shared memory");
#define MTG_LAUNCH_COLL(HN_)                                                                                   \
  {                                                                                                            \
    auto kern = aos ? collision_kernel<HN_, true> : collision_kernel<HN_, false>;                              \
    if (smem > 48 * 1024)                                                                                      \
      MTG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    kern<<<grid_dim, kCollWarps * 32, smem, s>>>(p);                                                           \
  }
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

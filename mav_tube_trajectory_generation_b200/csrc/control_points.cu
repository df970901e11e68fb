// libmtg_cuda.so — N2: Bezier control points of every segment and the reference's tube / end-cap / sphere
// constraints evaluated ON the control points (the conservative, convex-hull form of the corridor check).
//
// Replaces (reference, impl/polynomial_optimization_qcqp_impl.h = QC_I):
//   setupInverseControlPointMappingMatrix QC_I:267-319   B_inv(T): control points from endpoint derivatives
//   setupControlPointConstraints          QC_I:321-355   control point j of segment i = row j of B_inv_i times
//                                                        the segment's rows of C [d_f; d_p]
//   compute_tube_constraints              QC_I:369-429   |A x + b|^2 - r_tube^2 <= 0 on control points 1..N-2
//   compute_tube_end_constraints          QC_I:431-474   two half spaces on the same control points
//   compute_sphere_constraints            QC_I:357-365   |x - v_{i+1}|^2 - r^2 <= 0 on the last control point
// The reference only hands these to MOSEK as coefficients of d_p; here they are EVALUATED for given
// trajectories (a batch of candidates is screened in one launch). One thread per trajectory, segments in
// sequence; everything in registers (HN = N/2 is a template parameter).
//
// The forward matrix of QC_I:284-293 is lower triangular with the closed-form inverse
//   B_ul_inv(k, i) = binom(k, i) (n-i)!/n! T^i   (n = N - 1, i <= k),
// to which the reference's zeroing of entries in (-1e-5, 1e-5) (QC_I:300-306) is applied as written;
// B_lr_inv(h-1-k, i) = (-1)^i B_ul_inv(k, i) (QC_I:308-313).
#include "host_common.h"
#include "eval.cuh"  // TubeSeg, load_tube, at<AOS>()

MTG_REGISTER_BASE()

using namespace mtg;

namespace {

struct ControlPointParams {
  const double* __restrict__ coeffs;       // elem ((i*D + dim)*N + j), rec K*D*N; or nullptr
  const double* __restrict__ derivatives;  // elem ((v*h + k)*D + dim), rec (K+1)*h*D; or nullptr
  const double* __restrict__ seg_times;    // elem i, rec K
  const double* __restrict__ positions;    // elem v*3 + dim, rec (K+1)*3; or nullptr
  const double* __restrict__ radii;        // elem i*2 + {0,1}, rec K*2; or nullptr
  double* __restrict__ control_points;     // elem ((i*N + j)*D + dim), rec K*N*D; or nullptr
  double* __restrict__ tube;               // elem i*(N-2) + (j-1), rec K*(N-2); or nullptr
  double* __restrict__ cap_start;
  double* __restrict__ cap_end;
  double* __restrict__ sphere;             // elem i, rec K; or nullptr
  double* __restrict__ max_value;          // [B] or nullptr: largest constraint value of the trajectory
  uint8_t* __restrict__ feasible;          // [B] or nullptr: every value <= 0
  uint32_t* __restrict__ status;           // [B] or nullptr
  int B, b0, nb, K, D;
};

template <int HN, bool AOS>
__global__ void __launch_bounds__(128) control_points_kernel(const ControlPointParams p) {
  constexpr int N = 2 * HN, n = N - 1;
  const int local = blockIdx.x * blockDim.x + threadIdx.x;
  if (local >= p.nb) return;
  const int b = p.b0 + local;
  const size_t B = (size_t)p.B;
  const int K = p.K, D = p.D;
  const bool con = p.positions != nullptr && p.radii != nullptr && D == 3;
  const size_t rec_c = (size_t)K * D * N, rec_d = (size_t)(K + 1) * HN * D, rec_cp = (size_t)K * N * D;
  const size_t rec_q = (size_t)K * (N - 2);
  uint32_t st = 0;
  double worst = -INFINITY;
  // binom(k, i) (n-i)!/n!
  double w[HN][HN];
#pragma unroll
  for (int k = 0; k < HN; ++k) {
    double f = 1.0;  // (n-i)!/n!
#pragma unroll
    for (int i = 0; i < HN; ++i) {
      double bin = 1.0;
#pragma unroll
      for (int q = 0; q < i; ++q) bin = bin * (double)(k - q) / (double)(q + 1);
      w[k][i] = (i <= k) ? bin * f : 0.0;
      f = f / (double)(n - i);
    }
  }
  for (int i = 0; i < K; ++i) {
    double T = p.seg_times[at<AOS>((size_t)i, (size_t)K, B, b)];
    if (!(T > 0.0) || !(T < 1.7e308)) {  // LIN_I:296 CHECK_GT(segment_time, 0)
      st |= 1u;
      T = 1.0;
    }
    // B_ul_inv with the reference's zeroing
    double Bi[HN][HN];
    {
      double tp = 1.0;
#pragma unroll
      for (int c = 0; c < HN; ++c) {
#pragma unroll
        for (int k = 0; k < HN; ++k) {
          const double v = w[k][c] * tp;
          Bi[k][c] = (v > -0.00001 && v < 0.00001) ? 0.0 : v;
        }
        tp *= T;
      }
    }
    TubeSeg tg;
    double cp[N][4];
#pragma unroll
    for (int dim = 0; dim < 4; ++dim) {
      if (dim >= D) continue;
      double ds[HN], de[HN];
      if (p.derivatives) {
#pragma unroll
        for (int k = 0; k < HN; ++k) {
          ds[k] = p.derivatives[at<AOS>((size_t)((size_t)i * HN + k) * D + dim, rec_d, B, b)];
          de[k] = p.derivatives[at<AOS>((size_t)((size_t)(i + 1) * HN + k) * D + dim, rec_d, B, b)];
        }
      } else {
        double c[N];
#pragma unroll
        for (int j = 0; j < N; ++j) c[j] = p.coeffs[at<AOS>((size_t)(i * D + dim) * N + j, rec_c, B, b)];
#pragma unroll
        for (int k = 0; k < HN; ++k) {
          ds[k] = c_base.base[k * MTG_BASE_LD + k] * c[k];  // k! c_k
          double r = c_base.base[k * MTG_BASE_LD + n] * c[n];
#pragma unroll
          for (int j = n - 1; j >= 0; --j)
            if (j >= k) r = fma(r, T, c_base.base[k * MTG_BASE_LD + j] * c[j]);
          de[k] = r;
        }
      }
#pragma unroll
      for (int k = 0; k < HN; ++k) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int c = 0; c < HN; ++c) {
          s0 += Bi[k][c] * ds[c];
          s1 += ((c & 1) ? -Bi[k][c] : Bi[k][c]) * de[c];
        }
        cp[k][dim] = s0;
        cp[n - k][dim] = s1;
      }
      if (p.control_points) {
#pragma unroll
        for (int j = 0; j < N; ++j)
          p.control_points[at<AOS>((size_t)((size_t)i * N + j) * D + dim, rec_cp, B, b)] = cp[j][dim];
      }
    }
    if (!con) continue;
    // ---- constraints on the control points (feasible <=> value <= 0)
    {
      EvalParams ep = {};
      ep.positions = p.positions;
      ep.radii = p.radii;
      ep.B = p.B;
      ep.K = K;
      load_tube<AOS>(ep, i, b, tg);
    }
#pragma unroll
    for (int j = 1; j < N - 1; ++j) {
      const double x0 = cp[j][0], x1 = cp[j][1], x2 = cp[j][2];
      const double y0 = tg.A[0] * x0 + tg.A[1] * x1 + tg.A[2] * x2 + tg.bvec[0];
      const double y1 = tg.A[1] * x0 + tg.A[3] * x1 + tg.A[4] * x2 + tg.bvec[1];
      const double y2 = tg.A[2] * x0 + tg.A[4] * x1 + tg.A[5] * x2 + tg.bvec[2];
      const double vt = y0 * y0 + y1 * y1 + y2 * y2 - tg.r2;
      const double along = tg.n[0] * x0 + tg.n[1] * x1 + tg.n[2] * x2;
      const double vs = tg.cs - along;  // (-n).(x - p_start)
      const double ve = along - tg.ce;  //   n .(x - p_end)
      const size_t o = at<AOS>((size_t)i * (N - 2) + (j - 1), rec_q, B, b);
      if (p.tube) p.tube[o] = vt;
      if (p.cap_start) p.cap_start[o] = vs;
      if (p.cap_end) p.cap_end[o] = ve;
      worst = fmax(worst, fmax(vt, fmax(vs, ve)));
    }
    double vsph = -INFINITY;
    if (i < K - 1) {  // QC_I:349-351: no sphere on the last segment
      double sq = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double e = cp[n][k] - p.positions[at<AOS>((size_t)(i + 1) * 3 + k, (size_t)(K + 1) * 3, B, b)];
        sq += e * e;
      }
      const double rs = p.radii[at<AOS>((size_t)i * 2 + 1, (size_t)K * 2, B, b)];
      vsph = sq - rs * rs;
      worst = fmax(worst, vsph);
    }
    if (p.sphere) p.sphere[at<AOS>((size_t)i, (size_t)K, B, b)] = vsph;
  }
  if (p.max_value) p.max_value[b] = worst;
  if (p.feasible) p.feasible[b] = (uint8_t)((con && worst <= 0.0 && st == 0) ? 1 : 0);
  if (p.status) p.status[b] = st;
}

template <bool AOS>
int launch_cp(mtg_ctx* ctx, int N, const ControlPointParams& p, cudaStream_t s) {
  const int grid = (p.nb + 127) / 128;
  switch (N / 2) {
    case 1: control_points_kernel<1, AOS><<<grid, 128, 0, s>>>(p); break;
    case 2: control_points_kernel<2, AOS><<<grid, 128, 0, s>>>(p); break;
    case 3: control_points_kernel<3, AOS><<<grid, 128, 0, s>>>(p); break;
    case 4: control_points_kernel<4, AOS><<<grid, 128, 0, s>>>(p); break;
    case 5: control_points_kernel<5, AOS><<<grid, 128, 0, s>>>(p); break;
    case 6: control_points_kernel<6, AOS><<<grid, 128, 0, s>>>(p); break;
    default: return fail(ctx, MTG_ERR_UNSUPPORTED, "N must be even and <= 12");
  }
  ++ctx->launches;
  MTG_CUDA_TRY(cudaGetLastError());
  return MTG_OK;
}
}  // namespace

extern "C" int mtg_control_points_batch(mtg_ctx* ctx, const mtg_problem_desc* desc, const double* coeffs,
                                        const double* derivatives, const double* seg_times,
                                        const double* positions, const double* radii, double* control_points,
                                        double* tube, double* cap_start, double* cap_end, double* sphere,
                                        double* max_value, uint8_t* feasible, uint32_t* status, void* stream_) {
  int rc = validate_desc(ctx, desc);
  if (rc) return rc;
  if ((!coeffs && !derivatives) || !seg_times)
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "seg_times and one of coeffs / derivatives are required");
  const bool want_con = tube || cap_start || cap_end || sphere || max_value || feasible;
  if (want_con && (!positions || !radii || desc->D != 3))
    return fail(ctx, MTG_ERR_INVALID_ARGUMENT, "the constraints need positions, radii and D = 3 (QC_I:369-474)");
  if (desc->B == 0) return MTG_OK;
  MTG_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool aos = desc->layout == MTG_LAYOUT_AOS;
  const int B = desc->B, K = desc->K, D = desc->D, N = desc->N, h = N / 2;
  ControlPointParams p = {};
  p.K = K; p.D = D;
  auto launch = [&](cudaStream_t st) { return aos ? launch_cp<true>(ctx, N, p, st) : launch_cp<false>(ctx, N, p, st); };
  if (desc->memory == MTG_MEM_DEVICE) {
    p.coeffs = derivatives ? nullptr : coeffs; p.derivatives = derivatives; p.seg_times = seg_times;
    p.positions = want_con ? positions : nullptr; p.radii = want_con ? radii : nullptr;
    p.control_points = control_points; p.tube = tube; p.cap_start = cap_start; p.cap_end = cap_end;
    p.sphere = sphere; p.max_value = max_value; p.feasible = feasible; p.status = status;
    p.B = B; p.b0 = 0; p.nb = B;
    return launch(stream);
  }
  const size_t nq = (size_t)K * (N - 2);
  std::vector<HostTensor> ts = {
      {derivatives ? nullptr : coeffs, (size_t)K * D * N, 8, true, false, nullptr},
      {derivatives, (size_t)(K + 1) * h * D, 8, true, false, nullptr},
      {seg_times, (size_t)K, 8, true, false, nullptr},
      {want_con ? positions : nullptr, (size_t)(K + 1) * 3, 8, true, false, nullptr},
      {want_con ? radii : nullptr, (size_t)K * 2, 8, true, false, nullptr},
      {control_points, (size_t)K * N * D, 8, false, false, nullptr},
      {tube, nq, 8, false, false, nullptr}, {cap_start, nq, 8, false, false, nullptr},
      {cap_end, nq, 8, false, false, nullptr}, {sphere, (size_t)K, 8, false, false, nullptr},
      {max_value, 1, 8, false, true, nullptr}, {feasible, 1, 1, false, true, nullptr},
      {status, 1, 4, false, true, nullptr}};
  return run_chunked(ctx, stream, (size_t)B, aos, ts, [&](int nb, int C, cudaStream_t st) {
    p.coeffs = (const double*)ts[0].dev; p.derivatives = (const double*)ts[1].dev;
    p.seg_times = (const double*)ts[2].dev; p.positions = (const double*)ts[3].dev;
    p.radii = (const double*)ts[4].dev; p.control_points = (double*)ts[5].dev; p.tube = (double*)ts[6].dev;
    p.cap_start = (double*)ts[7].dev; p.cap_end = (double*)ts[8].dev; p.sphere = (double*)ts[9].dev;
    p.max_value = (double*)ts[10].dev; p.feasible = (uint8_t*)ts[11].dev; p.status = (uint32_t*)ts[12].dev;
    p.B = C; p.b0 = 0; p.nb = nb;
    return launch(st);
  });
}

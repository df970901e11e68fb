// The candidate sweep of INTEGRATION.md written against the C ABI alone, from compiled code with DEVICE memory:
// mtg_generate_candidates_batch -> mtg_solve_argmin_batch (fused running argmin, consecutive solves overlapping, the
// next batch generated one step ahead) -> the {cost, index} pair; checked against a host-side scan of the costs that
// mtg_solve_batch returns for the same candidates through HOST memory. Exit code 0 and "SWEEP OK" on success.
#include <cuda_runtime_api.h>
#include <mtg_cuda.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CHECK(call)                                                                         \
  do {                                                                                      \
    const int rc_ = (call);                                                                 \
    if (rc_ != 0) {                                                                         \
      std::printf("FAILED %s:%d  %s -> %d (%s)\n", __FILE__, __LINE__, #call, rc_, ctx ? mtg_last_error(ctx) : ""); \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

struct Pair {
  double cost;
  int64_t idx;
};

int main() {
  mtg_ctx* ctx = nullptr;
  CHECK(mtg_create(0, &ctx));
  const int B = 20000, K = 10, D = 3, N = 10, n_batches = 5;
  const uint64_t seed = 2024;
  mtg_problem_desc desc = {};
  desc.B = B;
  desc.K = K;
  desc.D = D;
  desc.N = N;
  desc.derivative_to_optimize = 4;
  desc.memory = MTG_MEM_DEVICE;
  desc.layout = MTG_LAYOUT_SOA;
  const double lo[3] = {-10, -10, -10}, hi[3] = {10, 10, 10};
  const size_t n_pos = (size_t)B * (K + 1) * D, n_t = (size_t)B * K;
  double *d_pos[2], *d_t[2];
  Pair* d_best = nullptr;
  for (int s = 0; s < 2; ++s) {
    CHECK(cudaMalloc((void**)&d_pos[s], n_pos * sizeof(double)));
    CHECK(cudaMalloc((void**)&d_t[s], n_t * sizeof(double)));
  }
  CHECK(cudaMalloc((void**)&d_best, sizeof(Pair)));
  cudaStream_t stream;
  CHECK(cudaStreamCreate(&stream));

  // ---- the sweep: device memory only, one launch per batch, nothing but the final pair comes back
  CHECK(mtg_set_solve_overlap(ctx, 1));
  CHECK(mtg_generate_candidates_batch(ctx, &desc, seed, 0, lo, hi, 3.0, 5.0, 6.5, d_pos[0], d_t[0], stream));
  for (int it = 0; it < n_batches; ++it) {
    if (it + 1 < n_batches)  // one batch ahead, into the other buffer pair (contract of mtg_set_solve_overlap)
      CHECK(mtg_generate_candidates_batch(ctx, &desc, seed, (int64_t)(it + 1) * B, lo, hi, 3.0, 5.0, 6.5,
                                          d_pos[(it + 1) & 1], d_t[(it + 1) & 1], stream));
    CHECK(mtg_solve_argmin_batch(ctx, &desc, d_pos[it & 1], nullptr, d_t[it & 1], nullptr, nullptr, nullptr, nullptr,
                                 (int64_t)it * B, it > 0, d_best, stream));
  }
  Pair got;
  CHECK(cudaMemcpyAsync(&got, d_best, sizeof(Pair), cudaMemcpyDeviceToHost, stream));
  CHECK(cudaStreamSynchronize(stream));
  CHECK(mtg_set_solve_overlap(ctx, 0));

  // ---- the same candidates through host memory, scanned serially like the reference's restart loop
  Pair want = {INFINITY, -1};
  std::vector<double> h_pos(n_pos), h_t(n_t), h_cost(B);
  std::vector<uint32_t> h_st(B);
  mtg_problem_desc hdesc = desc;
  hdesc.memory = MTG_MEM_HOST;
  for (int it = 0; it < n_batches; ++it) {
    CHECK(mtg_generate_candidates_batch(ctx, &desc, seed, (int64_t)it * B, lo, hi, 3.0, 5.0, 6.5, d_pos[0], d_t[0], stream));
    CHECK(cudaMemcpyAsync(h_pos.data(), d_pos[0], n_pos * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CHECK(cudaMemcpyAsync(h_t.data(), d_t[0], n_t * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CHECK(cudaStreamSynchronize(stream));
    CHECK(mtg_solve_batch(ctx, &hdesc, h_pos.data(), nullptr, h_t.data(), nullptr, h_cost.data(), nullptr, h_st.data(),
                          nullptr));
    for (int b = 0; b < B; ++b)
      if (h_st[b] == 0 && h_cost[b] == h_cost[b] && h_cost[b] < want.cost) {
        want.cost = h_cost[b];
        want.idx = (int64_t)it * B + b;
      }
  }
  std::printf("sweep of %d candidates: device pair {%.17g, %lld}, host scan {%.17g, %lld}\n", n_batches * B, got.cost,
              (long long)got.idx, want.cost, (long long)want.idx);
  const bool ok = got.cost == want.cost && got.idx == want.idx && got.idx >= 0;
  for (int s = 0; s < 2; ++s) {
    cudaFree(d_pos[s]);
    cudaFree(d_t[s]);
  }
  cudaFree(d_best);
  cudaStreamDestroy(stream);
  mtg_destroy(ctx);
  std::printf(ok ? "SWEEP OK\n" : "SWEEP FAILED\n");
  return ok ? 0 : 1;
}

// Write-pattern probe for the sweep kernels: B "trajectories", each owning a row region of S*24 bytes;
// one warp owns 16 trajectories and, chunk by chunk, writes PIECE bytes (default 768 = 32 samples x 24 B)
// to each of them as whole 256-byte warp stores - the store pattern of eval_tm_kernel without any
// compute or shared memory. Prints achieved write bandwidth per piece size and CTA residency.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/stream_probe tools/probes/stream_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(32) probe(double* out, int B, int S, int piece_doubles, int smem_pad) {
  extern __shared__ double pad[];
  const int lane = threadIdx.x;
  const int first = blockIdx.x * 16;
  if (first >= B) return;
  if (smem_pad < 0) pad[lane] = 0.0;  // never: keeps the allocation alive
  const size_t row = (size_t)S * 3;   // doubles per trajectory
  const int chunks = (int)(row / piece_doubles);
  for (int c = 0; c < chunks; ++c)
    for (int t = 0; t < 16; ++t) {
      double* dst = out + (size_t)(first + t) * row + (size_t)c * piece_doubles;
      for (int e = lane; e < piece_doubles; e += 32) dst[e] = (double)(c + t);
    }
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 262144, S = 1008;
  double* out;
  const size_t bytes = (size_t)B * S * 24;
  cudaMalloc(&out, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int pieces[] = {96, 192, 384, 1008};     // doubles: 768 B, 1536 B, 3072 B, 8064 B
  const int smems[] = {0, 16 * 1024, 24 * 1024, 36 * 1024};  // dynamic smem per CTA -> 32 (CTA cap), 13, 9, 6 warps/SM
  for (int sm : smems)
    for (int pd : pieces) {
      cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        probe<<<B / 16, 32, sm>>>(out, B, S, pd, 1);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const size_t written = (size_t)B * (size_t)((S * 3) / pd) * pd * 8;
      printf("{\"piece_bytes\": %d, \"smem_per_cta\": %d, \"ms\": %.3f, \"write_GBs\": %.1f}\n", pd * 8, sm, ms,
             written / (ms * 1e-3) / 1e9);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

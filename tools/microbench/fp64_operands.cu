// Does vector FP64 issue rate depend on operand distinctness / instruction kind?  (B200, sm_100a)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int MODE>
__global__ void k(double* out, const double* in, int iters) {
  constexpr int NA = 8;
  double acc[NA], x[NA], y[NA];
#pragma unroll
  for (int i = 0; i < NA; i++) { acc[i] = in[i]; x[i] = in[NA + i + (threadIdx.x & 1)]; y[i] = in[2 * NA + i + (threadIdx.x & 3)]; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NA; i++) {
      if (MODE == 0) acc[i] = fma(acc[i], x[0], y[0]);          // shared operands (reuse cache)
      if (MODE == 1) acc[i] = fma(x[i], y[i], acc[i]);          // 3 distinct registers per DFMA
      if (MODE == 2) acc[i] = acc[i] + x[i];                    // DADD distinct
      if (MODE == 3) acc[i] = acc[i] * x[i];                    // DMUL distinct
      if (MODE == 4) { double e = x[i] - acc[i]; acc[i] = fma(e, e, y[i]); }   // DADD + DFMA (distance pattern), dependent
      if (MODE == 5) { x[i] = y[i] - x[i]; acc[i] = fma(x[i], x[i], acc[i]); } // DADD + DFMA, accumulate
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NA; i++) s += acc[i] + x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(double* out, const double* in, int grid, int threads, int iters, int ops_per) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<MODE><<<grid, threads>>>(out, in, iters); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    CK(cudaEventRecord(e0)); k<MODE><<<grid, threads>>>(out, in, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  // warp-instructions per cycle per SM at 1.965 GHz
  double winst = (double)ops_per * 8 * iters * grid * (threads / 32);
  return winst / (best * 1e-3) / 148 / 1.965e9;
}

int main() {
  double *in, *out; CK(cudaMalloc(&in, 64 * 8)); CK(cudaMalloc(&out, 148 * 2 * 1024 * 8));
  double h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0 + 1e-9 * i; CK(cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice));
  const int it = 20000;
  for (int warps = 8; warps <= 32; warps *= 2) {
    int threads = warps * 32 > 1024 ? 1024 : warps * 32, grid = 148 * (warps * 32 / threads);
    printf("{\"warps_per_sm\": %d, \"unit\": \"FP64 warp-instr / cycle / SM\", \"dfma_shared_operands\": %.3f, \"dfma_distinct\": %.3f, "
           "\"dadd_distinct\": %.3f, \"dmul_distinct\": %.3f, \"dadd_dfma_dependent\": %.3f, \"dadd_dfma_accumulate\": %.3f}\n", warps,
           run<0>(out, in, grid, threads, it, 1), run<1>(out, in, grid, threads, it, 1), run<2>(out, in, grid, threads, it, 1),
           run<3>(out, in, grid, threads, it, 1), run<4>(out, in, grid, threads, it, 2), run<5>(out, in, grid, threads, it, 2));
  }
  return 0;
}

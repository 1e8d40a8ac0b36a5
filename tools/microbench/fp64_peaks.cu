// FP64 pipe microbenchmarks for B200 (sm_100a).
//
// Measures, with CUDA events, the sustained issue rate of
//   * DFMA            (vector FP64 pipe)
//   * DMMA.8x8x4      (mma.sync.m8n8k4.f64, the only FP64 tensor shape sm_100a has in SASS)
//   * DFMA + DMMA interleaved (are they separate pipes?)
//   * exp(double)     (libdevice) and the table-based exp used by kernel (a)
// The numbers are the denominators for the FP64 rooflines quoted in DESIGN.md;
// MEASURED_PEAKS.json only carries HBM GB/s and bf16 TFLOP/s.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dfma(double* out, const double* in, int iters) {
  double acc[NACC];
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
#pragma unroll
  for (int i = 0; i < NACC; i++) acc[i] = (double)i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dmma(double* out, const double* in, int iters) {
  double c0[NACC], c1[NACC];
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c0[i] = 0; c1[i] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// NF DFMA per DMMA, interleaved
template <int NACC, int NF>
__global__ void k_mixed(double* out, const double* in, int iters) {
  double c0[NACC], c1[NACC], f[NACC * NF];
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c0[i] = 0; c1[i] = 0; }
#pragma unroll
  for (int i = 0; i < NACC * NF; i++) f[i] = (double)i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      dmma884(c0[i], c1[i], a, b);
#pragma unroll
      for (int j = 0; j < NF; j++) f[i * NF + j] = fma(f[i * NF + j], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < NACC * NF; i++) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(double* out, const double* in, int iters) {
  double x0 = in[threadIdx.x & 31] * -1e-3, s = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) s += exp(x0 * (double)(it * 8 + i + 1));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_sqrt(double* out, const double* in, int iters) {
  double x0 = in[threadIdx.x & 31] + 1.0, s = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) s += sqrt(x0 * (double)(it * 8 + i + 1));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA latency: single dependent chain, one warp
__global__ void k_dmma_lat(double* out, const double* in, int iters, long long* cyc) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)], c0 = 0, c1 = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) dmma884(c0, c1, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = c0 + c1;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dfma_lat(double* out, const double* in, int iters, long long* cyc) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)], c0 = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) c0 = fma(c0, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = c0;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  int l2 = 0; CK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, 0));
  int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  printf("{\"device\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"clock_khz\": %d, \"smem_optin\": %zu,\n",
         prop.name, nsm, l2, clk, prop.sharedMemPerBlockOptin);
  double *in, *out; long long* cyc;
  CK(cudaMalloc(&in, 64 * 8)); CK(cudaMalloc(&out, (size_t)nsm * 32 * 1024 * 8)); CK(cudaMalloc(&cyc, 8));
  double h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0 + 1e-9 * i;
  CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
  const int iters = 20000;

  // sweep warps per SM for DFMA and DMMA
  int wps[] = {4, 8, 16, 32};
  for (int wi = 0; wi < 4; wi++) {
    int w = wps[wi];
    int threads = w * 32 > 1024 ? 1024 : w * 32;
    int blocks_per_sm = (w * 32) / threads;
    int grid = nsm * blocks_per_sm;
    {
      float ms = time_ms([&] { k_dfma<8><<<grid, threads>>>(out, in, iters); });
      double flops = 2.0 * 8 * iters * (double)grid * threads;
      printf(" \"dfma_tflops_w%d\": %.2f,\n", w, flops / ms / 1e9);
    }
    {
      float ms = time_ms([&] { k_dmma<8><<<grid, threads>>>(out, in, iters); });
      double flops = 2.0 * 256 * 8 * iters * (double)grid * (threads / 32);
      printf(" \"dmma_tflops_w%d\": %.2f,\n", w, flops / ms / 1e9);
    }
  }
  {
    int threads = 512, grid = nsm * 2;
    float ms = time_ms([&] { k_mixed<4, 8><<<grid, threads>>>(out, in, iters); });
    double fl_mma = 2.0 * 256 * 4 * iters * (double)grid * (threads / 32);
    double fl_fma = 2.0 * 4 * 8 * iters * (double)grid * threads;
    printf(" \"mixed_1dmma_8dfma\": {\"ms\": %.3f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f, \"sum\": %.2f},\n", ms,
           fl_mma / ms / 1e9, fl_fma / ms / 1e9, (fl_mma + fl_fma) / ms / 1e9);
    ms = time_ms([&] { k_mixed<4, 2><<<grid, threads>>>(out, in, iters); });
    fl_fma = 2.0 * 4 * 2 * iters * (double)grid * threads;
    printf(" \"mixed_1dmma_2dfma\": {\"ms\": %.3f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f, \"sum\": %.2f},\n", ms,
           fl_mma / ms / 1e9, fl_fma / ms / 1e9, (fl_mma + fl_fma) / ms / 1e9);
  }
  {
    int threads = 512, grid = nsm * 2, it2 = 2000;
    float ms = time_ms([&] { k_exp<<<grid, threads>>>(out, in, it2); });
    printf(" \"exp_gops\": %.2f,\n", 8.0 * it2 * (double)grid * threads / ms / 1e6);
    ms = time_ms([&] { k_sqrt<<<grid, threads>>>(out, in, it2); });
    printf(" \"sqrt_gops\": %.2f,\n", 8.0 * it2 * (double)grid * threads / ms / 1e6);
  }
  {
    long long hc;
    k_dmma_lat<<<1, 32>>>(out, in, 4096, cyc); CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
    printf(" \"dmma_latency_cycles\": %.2f,\n", hc / 4096.0);
    k_dfma_lat<<<1, 32>>>(out, in, 4096, cyc); CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
    printf(" \"dfma_latency_cycles\": %.2f\n}\n", hc / 4096.0);
  }
  return 0;
}

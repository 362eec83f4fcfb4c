// Microbenchmark: FP64 mma.sync.m8n8k4 (DMMA) vs DFMA issue rates on one B200.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k_dmma(double* out, int iters) {
  double acc[ILP][2];
#pragma unroll
  for (int u = 0; u < ILP; ++u) acc[u][0] = acc[u][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < ILP; ++u) dmma(acc[u][0], acc[u][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int u = 0; u < ILP; ++u) s += acc[u][0] + acc[u][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_dfma(double* out, int iters) {
  double acc[ILP];
#pragma unroll
  for (int u = 0; u < ILP; ++u) acc[u] = u;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < ILP; ++u) acc[u] = fma(acc[u], a, b);
  }
  double s = 0;
#pragma unroll
  for (int u = 0; u < ILP; ++u) s += acc[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: DMMA and DFMA streams in the same warp
template <int ILP>
__global__ void k_mix(double* out, int iters) {
  double acc[ILP][2], f[ILP];
#pragma unroll
  for (int u = 0; u < ILP; ++u) { acc[u][0] = acc[u][1] = 0.0; f[u] = u; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      dmma(acc[u][0], acc[u][1], a, b);
#pragma unroll
      for (int r = 0; r < 8; ++r) f[u] = fma(f[u], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int u = 0; u < ILP; ++u) s += acc[u][0] + acc[u][1] + f[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
float run(K kern, int blocks, int threads, double* out, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, threads>>>(out, iters);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  kern<<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    const int threads = warps * 32, blocks = sms;
    float m1 = run(k_dmma<1>, blocks, threads, out, iters);
    float m4 = run(k_dmma<4>, blocks, threads, out, iters);
    float f1 = run(k_dfma<1>, blocks, threads, out, iters);
    float f8 = run(k_dfma<8>, blocks, threads, out, iters);
    float x4 = run(k_mix<4>, blocks, threads, out, iters);
    auto tf = [&](double fma_per_thread_iter, float ms) { return 2.0 * fma_per_thread_iter * iters * threads * blocks / (ms * 1e-3) / 1e12; };
    // a DMMA = 8*8*4 = 256 FMA per warp = 8 per thread
    printf("warps/SM %2d: DMMA ilp1 %.2f TF (%.1f clk/inst/warp), ilp4 %.2f TF | DFMA ilp1 %.2f TF, ilp8 %.2f TF | mix(1 DMMA + 8 DFMA) ilp4 %.2f TF total\n",
           warps, tf(8, m1), m1 * 1e-3 * p.clockRate * 1e3 / iters, tf(8 * 4, m4), tf(1, f1), tf(8, f8), tf(4 * 16, x4));
  }
  return 0;
}

// Does the packed FP32 instruction (fma.rn.f32x2 -> FFMA2) issue at the FFMA rate on a B200, i.e. does it
// double the FP32 work per issue slot?  Chains of FFMA vs FFMA2 (and FMUL2 / FADD2), ILP 8, at 512 threads x 148 CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/probes/ffma2_probe.cu -o build_ab/ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_ffma(float* out, int iters) {
  float acc[ILP];
#pragma unroll
  for (int u = 0; u < ILP; ++u) acc[u] = u + threadIdx.x * 1e-6f;
  float a = 1.0f + threadIdx.x * 1e-7f, b = 1e-6f * (threadIdx.x + 1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < ILP; ++u) acc[u] = fmaf(acc[u], a, b);
  }
  float s = 0;
#pragma unroll
  for (int u = 0; u < ILP; ++u) s += acc[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP, int ILP>
__global__ void k_packed(float* out, int iters) {
  float2 acc[ILP];
#pragma unroll
  for (int u = 0; u < ILP; ++u) acc[u] = make_float2(u + threadIdx.x * 1e-6f, u + 0.5f);
  const float2 a = make_float2(1.0f + threadIdx.x * 1e-7f, 1.0f - threadIdx.x * 1e-7f);
  const float2 b = make_float2(1e-6f * (threadIdx.x + 1), 2e-6f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      if (OP == 0) acc[u] = __ffma2_rn(acc[u], a, b);
      if (OP == 1) acc[u] = __fmul2_rn(acc[u], a);
      if (OP == 2) acc[u] = __fadd2_rn(acc[u], b);
    }
  }
  float s = 0;
#pragma unroll
  for (int u = 0; u < ILP; ++u) s += acc[u].x + acc[u].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
static void run(const char* name, K kernel, int lanes_per_inst, int threads, float* out) {
  const int iters = 1 << 14, blocks = 148, ilp = 8;
  kernel<<<blocks, threads>>>(out, 16);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kernel<<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double inst = (double)iters * ilp * threads / 32 * blocks;  // warp instructions
  const double per_clk_sm = inst / (ms * 1e-3 * 1.965e9) / 148;
  printf("%-28s %4d thr  %8.3f ms  %5.2f warp-inst/clk/SM  %6.1f TFLOP/s (FMA = 2, MUL/ADD = 1)\n", name, threads, ms,
         per_clk_sm, inst * 32 * lanes_per_inst / (ms * 1e-3) / 1e12);
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  for (int threads : {256, 512, 1024}) {
    run("FFMA chain, ILP 8", k_ffma<8>, 2, threads, out);
    run("FFMA2 chain, ILP 8", k_packed<0, 8>, 4, threads, out);
    run("FMUL2 chain, ILP 8", k_packed<1, 8>, 2, threads, out);
    run("FADD2 chain, ILP 8", k_packed<2, 8>, 2, threads, out);
  }
  return 0;
}

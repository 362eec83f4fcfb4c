// Microbenchmarks for the gradient triple sweep on one B200 (round 2): what FP64 issue rate can the
// visit body reach (a) from registers, (b) with its stash loads from shared memory, at 8/12/16 warps
// per SM -- against the plain DFMA chain used as the roofline denominator.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I tad_dftd4_b200/csrc tools/probes/fp64_probe.cu -o build_ab/fp64_probe
#include <cstdio>
#include <cuda_runtime.h>

#include "d4b200_small.cuh"

using namespace d4b200;

#define KEEP(x) asm volatile("" : "+d"(x))

template <int ILP>
__global__ void k_dfma(double* out, int iters) {
  double acc[ILP];
#pragma unroll
  for (int u = 0; u < ILP; ++u) acc[u] = u + threadIdx.x * 1e-9;
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

// single-opcode chains: is DMUL / DADD issued at the DFMA rate?
template <int OP, int ILP>
__global__ void k_op(double* out, int iters) {
  double acc[ILP];
#pragma unroll
  for (int u = 0; u < ILP; ++u) acc[u] = 1.0 + u * 1e-3 + threadIdx.x * 1e-9;
  double a = 1.0 + 1e-12 * threadIdx.x, zero = 0.0, one = 1.0;
  KEEP(zero);
  KEEP(one);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      if (OP == 0) acc[u] = acc[u] * a;                                                      // DMUL
      if (OP == 1) acc[u] = acc[u] + a;                                                      // DADD
      if (OP == 2) asm("fma.rn.f64 %0, %1, %2, %3;" : "=d"(acc[u]) : "d"(acc[u]), "d"(a), "d"(zero));  // mul as FMA
      if (OP == 3) asm("fma.rn.f64 %0, %1, %2, %3;" : "=d"(acc[u]) : "d"(acc[u]), "d"(one), "d"(a));   // add as FMA
    }
  }
  double s = 0;
#pragma unroll
  for (int u = 0; u < ILP; ++u) s += acc[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mix of DFMA / DMUL / DADD in the proportions of the visit (3 distinct register operands per DFMA)
template <int ILP>
__global__ void k_mix(double* out, int iters) {
  double x[ILP], y[ILP], z[ILP];
#pragma unroll
  for (int u = 0; u < ILP; ++u) x[u] = 1.0 + u * 1e-3 + threadIdx.x * 1e-9, y[u] = 0.5 + u * 1e-3, z[u] = 1e-9 * u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      z[u] = fma(x[u], y[u], z[u]);
      x[u] = x[u] * y[u];
      y[u] = y[u] + z[u];
      z[u] = fma(y[u], x[u], z[u]);
    }
  }
  double s = 0;
#pragma unroll
  for (int u = 0; u < ILP; ++u) s += x[u] + y[u] + z[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the visit body, four visits per step, inputs in registers (opaque to the optimiser)
__global__ void k_visit_reg(double* out, int iters) {
  const double t = threadIdx.x * 1e-6;
  double a0 = 7.0 + t, a1 = 9.0 + t, c0 = 8.0 + t, c1 = 6.5 + t, P0 = 1e-3, P1 = 2e-3, Q0 = 1.5e-3, Q1 = 1.1e-3;
  double u0 = 0.3, u1 = 0.2, v0 = 0.25, v1 = 0.35;
  OwnerPair<double> o[4];
  for (int k = 0; k < 4; ++k) {
    o[k].b = 5.0 + k + t, o[k].b2 = o[k].b * o[k].b, o[k].twob = 2 * o[k].b, o[k].iP = 900.0 + k, o[k].sPu = 1500.0 + k,
    o[k].kAi = 2400.0 + k;
  }
  double g[4] = {0, 0, 0, 0}, c[4] = {0, 0, 0, 0}, s[4] = {0, 0, 0, 0}, dH = 0, dL = 0;
  const double z = 0.0, kB = 0.1666;
  for (int it = 0; it < iters; ++it) {
    KEEP(a0); KEEP(a1); KEEP(c0); KEEP(c1); KEEP(P0); KEEP(P1); KEEP(Q0); KEEP(Q1); KEEP(u0); KEEP(u1); KEEP(v0); KEEP(v1);
    grad_visit<double, false, true>(a0, P0, u0, c0, Q0, v0, o[0].b, o[0].b2, o[0].twob, z, o[0].iP, o[0].sPu, o[0].kAi, kB, z, z, z, z, g[0], c[0], s[0], dH, dL);
    grad_visit<double, false, true>(a0, P0, u0, c1, Q1, v1, o[1].b, o[1].b2, o[1].twob, z, o[1].iP, o[1].sPu, o[1].kAi, kB, z, z, z, z, g[1], c[1], s[1], dH, dL);
    grad_visit<double, false, true>(a1, P1, u1, c0, Q0, v0, o[2].b, o[2].b2, o[2].twob, z, o[2].iP, o[2].sPu, o[2].kAi, kB, z, z, z, z, g[2], c[2], s[2], dH, dL);
    grad_visit<double, false, true>(a1, P1, u1, c1, Q1, v1, o[3].b, o[3].b2, o[3].twob, z, o[3].iP, o[3].sPu, o[3].kAi, kB, z, z, z, z, g[3], c[3], s[3], dH, dL);
  }
  double r = 0;
  for (int k = 0; k < 4; ++k) r += g[k] + c[k] + s[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// same, the twelve stash values of a step loaded from shared memory (4 broadcast rows + unit-stride columns)
template <int UNROLL>
__global__ void k_visit_lds(double* out, int iters, int n) {
  extern __shared__ double sm[];  // three planes [n][64]
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 3 * n * 64; i += blockDim.x) sm[i] = 5.0 + (i % 97) * 0.01;
  __syncthreads();
  const double t = threadIdx.x * 1e-6;
  OwnerPair<double> o[4];
  for (int k = 0; k < 4; ++k) {
    o[k].b = 5.0 + k + t, o[k].b2 = o[k].b * o[k].b, o[k].twob = 2 * o[k].b, o[k].iP = 900.0 + k, o[k].sPu = 1500.0 + k,
    o[k].kAi = 2400.0 + k;
  }
  double g[4] = {0, 0, 0, 0}, c[4] = {0, 0, 0, 0}, s[4] = {0, 0, 0, 0}, dH = 0, dL = 0;
  const double z = 0.0, kB = 0.1666;
  const double *pa = sm, *pP = sm + n * 64, *pu = sm + 2 * n * 64;
  for (int it = 0; it < iters; ++it) {
#pragma unroll UNROLL
    for (int j = 0; j < n; ++j) {
      const int xi0 = j * 64 + 62, xi1 = j * 64 + 63, xk0 = j * 64 + lane, xk1 = j * 64 + 32 - 2 + lane;
      const double a0 = pa[xi0], P0 = pP[xi0], u0 = pu[xi0], a1 = pa[xi1], P1 = pP[xi1], u1 = pu[xi1];
      const double c0 = pa[xk0], Q0 = pP[xk0], v0 = pu[xk0], c1 = pa[xk1], Q1 = pP[xk1], v1 = pu[xk1];
      grad_visit<double, false, true>(a0, P0, u0, c0, Q0, v0, o[0].b, o[0].b2, o[0].twob, z, o[0].iP, o[0].sPu, o[0].kAi, kB, z, z, z, z, g[0], c[0], s[0], dH, dL);
      grad_visit<double, false, true>(a0, P0, u0, c1, Q1, v1, o[1].b, o[1].b2, o[1].twob, z, o[1].iP, o[1].sPu, o[1].kAi, kB, z, z, z, z, g[1], c[1], s[1], dH, dL);
      grad_visit<double, false, true>(a1, P1, u1, c0, Q0, v0, o[2].b, o[2].b2, o[2].twob, z, o[2].iP, o[2].sPu, o[2].kAi, kB, z, z, z, z, g[2], c[2], s[2], dH, dL);
      grad_visit<double, false, true>(a1, P1, u1, c1, Q1, v1, o[3].b, o[3].b2, o[3].twob, z, o[3].iP, o[3].sPu, o[3].kAi, kB, z, z, z, z, g[3], c[3], s[3], dH, dL);
    }
  }
  double r = 0;
  for (int k = 0; k < 4; ++k) r += g[k] + c[k] + s[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F>
static float timeit(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  return best;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
  const double peak_lanes = 64.0;  // FP64 lanes per SM per clock
  printf("SMs %d; utilisation = FP64 thread-instructions / (time x 64 lanes x SMs x 1.965 GHz)\n", sms);
  auto report = [&](const char* name, double fp64_inst_per_thread, int blocks, int threads, float ms) {
    const double inst = fp64_inst_per_thread * blocks * threads;
    printf("%-44s %4d thr x %4d blk  %8.3f ms  FP64 issue utilisation %5.1f %%\n", name, threads, blocks, ms,
           100.0 * inst / (ms * 1e-3 * peak_lanes * sms * 1.965e9));
  };
  const int it = 1 << 14;
  for (int th : {256, 512, 1024}) {
    float ms = timeit([&] { k_dfma<8><<<sms, th>>>(out, it); });
    report("DFMA chain, ILP 8", 8.0 * it, sms, th, ms);
    ms = timeit([&] { k_mix<4><<<sms, th>>>(out, it); });
    report("DFMA/DMUL/DADD mix (2:1:1), ILP 4", 16.0 * it, sms, th, ms);
  }
  for (int th : {512}) {
    float ms = timeit([&] { k_op<0, 8><<<sms, th>>>(out, it); });
    report("DMUL chain, ILP 8", 8.0 * it, sms, th, ms);
    ms = timeit([&] { k_op<1, 8><<<sms, th>>>(out, it); });
    report("DADD chain, ILP 8", 8.0 * it, sms, th, ms);
    ms = timeit([&] { k_op<2, 8><<<sms, th>>>(out, it); });
    report("multiply as fma(a, b, 0), ILP 8", 8.0 * it, sms, th, ms);
    ms = timeit([&] { k_op<3, 8><<<sms, th>>>(out, it); });
    report("add as fma(a, 1, b), ILP 8", 8.0 * it, sms, th, ms);
  }
  const int vit = 1 << 12;
  for (int th : {256, 384, 512}) {
    float ms = timeit([&] { k_visit_reg<<<sms, th>>>(out, vit); });
    report("visit body x4, inputs in registers", 84.0 * vit, sms, th, ms);
  }
  cudaFuncSetAttribute(k_visit_lds<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_visit_lds<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int n = 100, lit = 40;
  for (int th : {256, 384, 512}) {
    float ms = timeit([&] { k_visit_lds<1><<<sms, th, 3 * n * 64 * 8>>>(out, lit, n); });
    report("visit body x4, stash from shared memory", 84.0 * lit * n, sms, th, ms);
    ms = timeit([&] { k_visit_lds<2><<<sms, th, 3 * n * 64 * 8>>>(out, lit, n); });
    report("  same, two steps in flight", 84.0 * lit * n, sms, th, ms);
  }
  return 0;
}

// Kernel flavours = (real type, energy|gradient, D4|D4S) x five size classes.  Each
// flavour lives in its own translation unit (flavour_*.cu) so that the 40 kernel
// instantiations compile in parallel; the API unit only sees these two functions.
#pragma once

#include <cuda_runtime.h>

#include "d4b200_small_args.cuh"

struct d4b200_tables;

// Size classes per flavour: X(class, CAP, threads, min CTAs/SM for launch bounds).
// Caps are bounded by the 227 KB shared-memory budget (Lay<>::total, static_assert).
#define D4_CLASSES_F64_E(X) X(0, 32, 128, 6) X(1, 48, 192, 4) X(2, 64, 256, 3) X(3, 96, 512, 1) X(4, 120, 512, 1)
#ifndef D4_G100_NT
#define D4_G100_NT 512  // threads of the largest FP64 gradient class (A/B knob)
#endif
#define D4_CLASSES_F64_G(X) X(0, 32, 128, 4) X(1, 48, 256, 2) X(2, 64, 512, 1) X(3, 80, 512, 1) X(4, 100, D4_G100_NT, 1)
#define D4_CLASSES_F32_E(X) X(0, 32, 128, 6) X(1, 48, 192, 4) X(2, 64, 256, 4) X(3, 96, 512, 2) X(4, 128, 512, 1)
#define D4_CLASSES_F32_G(X) X(0, 32, 128, 6) X(1, 48, 256, 3) X(2, 64, 512, 2) X(3, 96, 512, 1) X(4, 128, 512, 1)
// D4S: no per-atom polarizability vectors (pair-dependent weights are evaluated per
// pair), but the weights cannot alias the planes -> slightly different caps
#define D4_CLASSES_F64_E_S(X) X(0, 32, 128, 4) X(1, 48, 192, 3) X(2, 64, 256, 3) X(3, 96, 512, 1) X(4, 120, 512, 1)
#define D4_CLASSES_F64_G_S(X) X(0, 32, 128, 4) X(1, 48, 256, 2) X(2, 64, 512, 1) X(3, 96, 512, 1) X(4, 120, 512, 1)
#define D4_CLASSES_F32_E_S(X) X(0, 32, 128, 4) X(1, 48, 192, 3) X(2, 64, 256, 3) X(3, 96, 512, 1) X(4, 128, 512, 1)
#define D4_CLASSES_F32_G_S(X) X(0, 32, 128, 4) X(1, 48, 256, 2) X(2, 64, 512, 1) X(3, 96, 512, 1) X(4, 128, 512, 1)

#define D4_DECLARE(NAME, TYPE)                 \
  int d4_configure_##NAME(d4b200_tables* h);   \
  void d4_launch_##NAME(int c, unsigned grid, cudaStream_t st, const d4b200::SmallArgs<TYPE>& A);
D4_DECLARE(f64_e, double)
D4_DECLARE(f64_g, double)
D4_DECLARE(f32_e, float)
D4_DECLARE(f32_g, float)
D4_DECLARE(f64_e_s, double)
D4_DECLARE(f64_g_s, double)
D4_DECLARE(f32_e_s, float)
D4_DECLARE(f32_g_s, float)
#undef D4_DECLARE

template <typename T>
inline int flavour_configure(d4b200_tables* h, bool grad, int model);
template <>
inline int flavour_configure<double>(d4b200_tables* h, bool grad, int model) {
  if (model == 0) return grad ? d4_configure_f64_g(h) : d4_configure_f64_e(h);
  return grad ? d4_configure_f64_g_s(h) : d4_configure_f64_e_s(h);
}
template <>
inline int flavour_configure<float>(d4b200_tables* h, bool grad, int model) {
  if (model == 0) return grad ? d4_configure_f32_g(h) : d4_configure_f32_e(h);
  return grad ? d4_configure_f32_g_s(h) : d4_configure_f32_e_s(h);
}

inline void flavour_launch_impl(bool grad, int model, int c, unsigned grid, cudaStream_t st,
                                const d4b200::SmallArgs<double>& A) {
  if (model == 0) return grad ? d4_launch_f64_g(c, grid, st, A) : d4_launch_f64_e(c, grid, st, A);
  return grad ? d4_launch_f64_g_s(c, grid, st, A) : d4_launch_f64_e_s(c, grid, st, A);
}
inline void flavour_launch_impl(bool grad, int model, int c, unsigned grid, cudaStream_t st,
                                const d4b200::SmallArgs<float>& A) {
  if (model == 0) return grad ? d4_launch_f32_g(c, grid, st, A) : d4_launch_f32_e(c, grid, st, A);
  return grad ? d4_launch_f32_g_s(c, grid, st, A) : d4_launch_f32_e_s(c, grid, st, A);
}
template <typename T>
inline void flavour_launch(bool grad, int model, int c, unsigned grid, cudaStream_t st,
                           const d4b200::SmallArgs<T>& A) {
  flavour_launch_impl(grad, model, c, grid, st, A);
}

// ---- definition helper used by the flavour_*.cu translation units ----------
#define D4_CFG_ONE(C, CAPV, NTV, MINBV)                                                            \
  {                                                                                                \
    using LL = d4b200::Lay<D4_TYPE, D4_GRAD, D4_S, CAPV>;                                          \
    static_assert(LL::total <= 227 * 1024, "class does not fit the shared-memory budget");         \
    auto kern = d4b200::small_kernel<D4_TYPE, D4_GRAD, D4_S, CAPV, NTV, MINBV>;                    \
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LL::total);   \
    if (e != cudaSuccess) return (int)e;                                                           \
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NTV, LL::total);                 \
    if (e != cudaSuccess) return (int)e;                                                           \
    if (occ < 1) occ = 1;                                                                          \
    if (occ > class_occ_cap(C)) occ = class_occ_cap(C);                                            \
    h->caps[md][dt][gr][C] = CAPV;                                                                 \
    h->threads[md][dt][gr][C] = NTV;                                                               \
    h->smem[md][dt][gr][C] = LL::total;                                                            \
    h->grid_per_sm[md][dt][gr][C] = occ;                                                           \
  }
#define D4_LAUNCH_ONE(C, CAPV, NTV, MINBV)                                                         \
  case C:                                                                                          \
    d4b200::small_kernel<D4_TYPE, D4_GRAD, D4_S, CAPV, NTV, MINBV>                                 \
        <<<grid, NTV, d4b200::Lay<D4_TYPE, D4_GRAD, D4_S, CAPV>::total, st>>>(A);                  \
    break;
#define D4_DEFINE_FLAVOUR(NAME, LIST)                                                              \
  int d4_configure_##NAME(d4b200_tables* h) {                                                      \
    constexpr int dt = sizeof(D4_TYPE) == 8 ? 0 : 1, gr = D4_GRAD ? 1 : 0, md = D4_S ? 1 : 0;      \
    cudaError_t e = cudaSuccess;                                                                   \
    int occ = 0;                                                                                   \
    LIST(D4_CFG_ONE)                                                                               \
    return 0;                                                                                      \
  }                                                                                                \
  void d4_launch_##NAME(int c, unsigned grid, cudaStream_t st,                                     \
                        const d4b200::SmallArgs<D4_TYPE>& A) {                                     \
    switch (c) { LIST(D4_LAUNCH_ONE) }                                                             \
  }

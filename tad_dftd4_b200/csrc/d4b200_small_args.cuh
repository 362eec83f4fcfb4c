// Kernel argument block of the one-CTA-per-structure kernels.
#pragma once

#include "d4b200_common.cuh"

namespace d4b200 {

template <typename T>
struct SmallArgs {
  const int64_t* numbers;
  const T* pos;
  const T* q;
  const T* gin;  // upstream dL/dE (nullable = ones)
  T* energy;
  T* cn_out;
  T* grad;
  T* gradq;
  T* c6_out;     // properties mode: [nbatch, nat, nat] pair C6 (pre-zeroed by the host side)
  T* alpha_out;  // properties mode: [nbatch, nat] static polarizabilities
  T* scratch;  // [gridDim.x][2 or 3][CAP(CAP-1)/2] per-CTA, L2-resident per-pair results
  int nbatch, nat, cls;
  unsigned long long* phase;  // optional [16] per-phase cycle counters (development profiling)
  Tables<T> tab;
  Par<T> par;
  Work wk;
};

}  // namespace d4b200

// Shared definitions for the sm_100a DFT-D4 kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/d4b200.h"

namespace d4b200 {

constexpr int NELEM = 104;  // reference tables: Z = 0 (dummy) .. 103
constexpr int NREF = 7;
constexpr int NFREQ = 23;
constexpr int SMALL_MAX = 128;  // largest structure of the one-CTA-per-structure family
constexpr int NCLASS = 5;       // size classes of the small family
constexpr int HIST_BINS = SMALL_MAX + 2;
constexpr int D4S_ECAP = 8;         // D4S weight table: distinct elements per structure it can hold
constexpr int D4S_WSTR = 2 * NREF;  // table entry: gw[7], d gw/d cn [7]
constexpr int D4S_RCAP = 7;         // D4S gradient: distinct elements whose reference-C6 blocks are staged in shared memory

// Per-element tables in device memory (layout: tad_dftd4_b200/tables.py).
// Weight-related tables are always double (the reference evaluates the
// Gaussian weights in float64 for every dtype, model/d4.py:162-207).
template <typename T>
struct Tables {
  const T* rcov;       // [NELEM]
  const T* r4r2;       // [NELEM]
  const T* sqrt_r4r2;  // [NELEM] 3^(1/4) sqrt(r4r2)
  const T* den;        // [NELEM*NELEM] CN electronegativity factor
  const T* alpha_w;    // [NELEM*NREF*NFREQ] sqrt(3/pi w) * alpha
  const T* alpha0;     // [NELEM*NREF] alpha(i0)
  const T* rc6;        // [NELEM*NELEM*NREF*NREF] reference C6 (D4S pair contraction)
  const double* gamgc;   // [NELEM]
  const double* zeff;    // [NELEM]
  const double* refcn;   // [NELEM*NREF]
  const double* refq;    // [NELEM*NREF]
  const double* zeta0;   // [NELEM*NREF]
  const double* wfpair;  // [NELEM*NELEM]
  const int* refc;       // [NELEM*NREF]
  const int* maxcn_ref;  // [NELEM]
  const unsigned short* pij;  // [SMALL_MAX*(SMALL_MAX-1)/2] pair index p -> (hi << 8) | lo
};

// Kernel-side view of d4b200_params in the compute type.
template <typename T>
struct Par {
  T s6, s8, s10k /* s10*49/40 */, a1, a2, alp3 /* alp/3 */;
  T fac9;  /* cbrt(s9/6) folded into the ATM pair factor */
  T disp2_sq, disp3_sq, cn_sq;
  double wf, ga;
  int has_atm;  // s9 != 0
  int alp16;    // alp == 16 (default): (R0/r)^(16/3) = x^5 cbrt(x)
  int model;
};

// Integer workspace layout (device).
struct Work {
  int* nreal;        // [nbatch]
  int* order;        // [nbatch] structures sorted by descending size
  int* hist;         // [HIST_BINS] (bin SMALL_MAX+1 = too large)
  int* cursor;       // [HIST_BINS]
  int* class_range;  // [2*(NCLASS+1)]  begin/end in `order`; slot NCLASS = too large
  int* queue;        // [NCLASS+1] dynamic work counters
  int* status;       // [1]
  int* done;         // [1] blocks of the planning kernel that have finished counting
};

__host__ __device__ inline int tri(int hi, int lo) { return hi * (hi - 1) / 2 + lo; }

}  // namespace d4b200

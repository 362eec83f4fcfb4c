// Device-table handle shared by the API translation unit and the per-flavour
// kernel translation units.
#pragma once

#include <cuda_runtime.h>

#include "d4b200_common.cuh"

using namespace d4b200;

struct d4b200_tables {
  int device;
  int num_sms;
  double ga, gc;
  double* f64;  // device copy of the double blob
  float* f32;   // same blob converted to float
  int* i32;
  size_t n_f64, n_i32;
  Tables<double> t64;
  Tables<float> t32;
  // cached launch configuration: [model][dtype][grad][class]
  int caps[2][2][2][NCLASS];
  int threads[2][2][2][NCLASS];
  int grid_per_sm[2][2][2][NCLASS];
  size_t smem[2][2][2][NCLASS];
  // classes are independent: they run concurrently on a small stream pool
  cudaStream_t cstream[NCLASS];
  cudaEvent_t ev_fork, ev_join[NCLASS];
  // optional per-launch timing (bench.py roofline): events around each class kernel
  int profile;
  cudaEvent_t ev[2 * NCLASS];
  cudaEvent_t ev_call[3];  // call start, prep done, call end
  unsigned long long* phase_dev;  // [NCLASS][16] per-phase cycle counters (development)
  int phase_on;
  int ev_used[NCLASS];
};

// upper bound of resident CTAs per SM we ever launch for a class
inline int class_occ_cap(int c) { return c == 0 ? 10 : c == 1 ? 5 : c == 2 ? 4 : 2; }

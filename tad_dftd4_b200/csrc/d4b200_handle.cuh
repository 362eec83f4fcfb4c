// Device-table handle shared by the API translation unit and the per-flavour
// kernel translation units.
#pragma once

#include <cuda_runtime.h>

#include "d4b200_common.cuh"

using namespace d4b200;

#define D4_HOST_SLOTS 4
#define D4_HOST_STATUS 64  // chunks per host-buffer call whose status words are read back

struct d4b200_tables {
  int device;
  int num_sms;
  double ga, gc;
  double* f64;  // device copy of the double blob
  float* f32;   // same blob converted to float
  int* i32;
  unsigned short* pij;  // pair index table of the small family
  size_t n_f64, n_i32;
  Tables<double> t64;
  Tables<float> t32;
  // cached launch configuration: [model][dtype][grad][class]
  int caps[2][2][2][NCLASS];
  int threads[2][2][2][NCLASS];
  int grid_per_sm[2][2][2][NCLASS];
  size_t smem[2][2][2][NCLASS];
  // classes are independent: they run concurrently on a small stream pool
  // class stream pools: pool 0 serves calls on the caller's stream, pool 1+s the host slot s
  cudaStream_t cstream[1 + D4_HOST_SLOTS][NCLASS];
  cudaEvent_t ev_fork[1 + D4_HOST_SLOTS], ev_join[1 + D4_HOST_SLOTS][NCLASS];
  // optional per-launch timing (bench.py roofline): events around each class kernel
  int profile;
  cudaEvent_t ev[2 * NCLASS];
  cudaEvent_t ev_call[3];  // call start, prep done, call end
  unsigned long long* phase_dev;  // [NCLASS][16] per-phase cycle counters (development)
  int phase_on;
  int ev_used[NCLASS];
  // host-buffer entry points: D4_HOST_SLOTS pipeline slots (stream, staging buffers, workspace)
  cudaStream_t hstream[D4_HOST_SLOTS];
  void* hbuf[D4_HOST_SLOTS];
  size_t hbuf_bytes[D4_HOST_SLOTS];
  cudaStream_t hcopy;  // all H2D copies, in chunk order
  cudaEvent_t hev_in[D4_HOST_SLOTS], hev_done[D4_HOST_SLOTS];
  int* hstatus;  // pinned host: status word of every chunk of the last host-buffer call
};

// upper bound of resident CTAs per SM we ever launch for a class
inline int class_occ_cap(int c) { return c == 0 ? 10 : c == 1 ? 5 : c == 2 ? 4 : 2; }

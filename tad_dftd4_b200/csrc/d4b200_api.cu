// C-ABI entry points (include/d4b200.h) + batch preparation kernels.
#include <cuda_runtime.h>
#include <cstdlib>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "d4b200_common.cuh"
#include "d4b200_flavour.cuh"
#include "d4b200_handle.cuh"
#include "d4b200_small_args.cuh"



static thread_local int g_launches = 0;
static long long g_total_launches = 0;  // cumulative, all calls of this process

namespace {

// section offsets inside the double blob (must match tables.py F64_LAYOUT)
struct BlobOffsets {
  size_t rcov, r4r2, sqrt_r4r2, gamgc, zeff, refcn, refq, zeta0, alpha0, den, alpha_w, wfpair, rc6, total;
  size_t refc, maxcn_ref, itotal;
};
BlobOffsets blob_offsets() {
  BlobOffsets o;
  size_t p = 0;
  o.rcov = p, p += NELEM;
  o.r4r2 = p, p += NELEM;
  o.sqrt_r4r2 = p, p += NELEM;
  o.gamgc = p, p += NELEM;
  o.zeff = p, p += NELEM;
  o.refcn = p, p += NELEM * NREF;
  o.refq = p, p += NELEM * NREF;
  o.zeta0 = p, p += NELEM * NREF;
  o.alpha0 = p, p += NELEM * NREF;
  o.den = p, p += NELEM * NELEM;
  o.alpha_w = p, p += NELEM * NREF * NFREQ;
  o.wfpair = p, p += NELEM * NELEM;
  o.rc6 = p, p += (size_t)NELEM * NELEM * NREF * NREF;
  o.total = p;
  size_t q = 0;
  o.refc = q, q += NELEM * NREF;
  o.maxcn_ref = q, q += NELEM;
  o.itotal = q;
  return o;
}

template <typename T>
Tables<T> make_tables(const T* real, const double* f64, const int* i32, const unsigned short* pij) {
  const BlobOffsets o = blob_offsets();
  Tables<T> t;
  t.rcov = real + o.rcov;
  t.r4r2 = real + o.r4r2;
  t.sqrt_r4r2 = real + o.sqrt_r4r2;
  t.den = real + o.den;
  t.alpha_w = real + o.alpha_w;
  t.alpha0 = real + o.alpha0;
  t.rc6 = real + o.rc6;
  t.gamgc = f64 + o.gamgc;
  t.zeff = f64 + o.zeff;
  t.refcn = f64 + o.refcn;
  t.refq = f64 + o.refq;
  t.zeta0 = f64 + o.zeta0;
  t.wfpair = f64 + o.wfpair;
  t.refc = i32 + o.refc;
  t.maxcn_ref = i32 + o.maxcn_ref;
  t.pij = pij;
  return t;
}

// p = hi (hi - 1) / 2 + lo  ->  (hi << 8) | lo, for every pair of a SMALL_MAX-atom structure
__global__ void k_pair_table(unsigned short* __restrict__ pij) {
  const int hi = blockIdx.x + 1;
  for (int lo = threadIdx.x; lo < hi; lo += blockDim.x) pij[hi * (hi - 1) / 2 + lo] = (unsigned short)((hi << 8) | lo);
}

__global__ void k_to_float(const double* __restrict__ in, float* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

// ---- batch preparation: size histogram -> descending-size order ------------
__global__ void k_count(const int64_t* __restrict__ numbers, int nbatch, int nat, Work wk) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= nbatch) return;
  const int64_t* row = numbers + (size_t)warp * nat;
  int c = 0;
  for (int t = lane; t < nat; t += 32) c += row[t] != 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) {
    wk.nreal[warp] = c;
    atomicAdd(&wk.hist[c <= SMALL_MAX ? c : SMALL_MAX + 1], 1);
  }
}

struct Caps {
  int v[NCLASS];
};

__global__ void k_scan(int nbatch, Work wk, Caps caps) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  // start[n] = number of structures with more than n atoms
  int pos = 0;
  for (int n = SMALL_MAX + 1; n >= 0; --n) {
    wk.cursor[n] = pos;
    pos += wk.hist[n];
  }
  // after the loop cursor[n] = start[n]; class c owns sizes (caps[c-1], caps[c]]
  int end = nbatch;
  for (int c = 0; c < NCLASS; ++c) {
    const int begin = wk.cursor[caps.v[c]];
    wk.class_range[2 * c] = begin;
    wk.class_range[2 * c + 1] = end;
    end = begin;
  }
  wk.class_range[2 * NCLASS] = 0;  // too large for the small family
  wk.class_range[2 * NCLASS + 1] = end;
  if (end > 0) atomicOr(wk.status, D4B200_STATUS_TOO_LARGE);
}

__global__ void k_scatter(int nbatch, Work wk) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbatch) return;
  const int n = wk.nreal[b];
  const int pos = atomicAdd(&wk.cursor[n <= SMALL_MAX ? n : SMALL_MAX + 1], 1);
  wk.order[pos] = b;
}

// The three steps above in ONE launch for batches of moderate size: every block counts
// the atoms of its structures; the block that finishes last (ticket counter) turns the
// histogram into class ranges and scatters the structures into `order`.
constexpr int PLAN_THREADS = 1024;
constexpr int PLAN_MAX_BATCH = 32768;

__global__ void __launch_bounds__(PLAN_THREADS) k_plan(const int64_t* __restrict__ numbers, int nbatch, int nat,
                                                        Work wk, Caps caps) {
  __shared__ int s_cursor[HIST_BINS];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31;
  const int b = (blockIdx.x * PLAN_THREADS + tid) >> 5;
  if (b < nbatch) {
    const int64_t* row = numbers + (size_t)b * nat;
    int c = 0;
    for (int t = lane; t < nat; t += 32) c += row[t] != 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) {
      wk.nreal[b] = c;
      atomicAdd(&wk.hist[c <= SMALL_MAX ? c : SMALL_MAX + 1], 1);
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(wk.done, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < HIST_BINS) s_cursor[tid] = __ldcg(&wk.hist[tid]);
  __syncthreads();
  if (tid == 0) {
    int pos = 0;  // start[n] = number of structures with more than n atoms
    for (int n = SMALL_MAX + 1; n >= 0; --n) {
      const int h = s_cursor[n];
      s_cursor[n] = pos;
      pos += h;
    }
    int end = nbatch;
    for (int c = 0; c < NCLASS; ++c) {
      const int begin = s_cursor[caps.v[c]];
      wk.class_range[2 * c] = begin;
      wk.class_range[2 * c + 1] = end;
      end = begin;
    }
    wk.class_range[2 * NCLASS] = 0;  // too large for the small family
    wk.class_range[2 * NCLASS + 1] = end;
    if (end > 0) atomicOr(wk.status, D4B200_STATUS_TOO_LARGE);
  }
  __syncthreads();
  for (int s = tid; s < nbatch; s += PLAN_THREADS) {
    const int n = __ldcg(&wk.nreal[s]);
    const int pos = atomicAdd(&s_cursor[n <= SMALL_MAX ? n : SMALL_MAX + 1], 1);
    wk.order[pos] = s;
  }
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int HEADER_INTS = 2 + (NCLASS + 1) + 2 * (NCLASS + 1) + 2 * HIST_BINS;

size_t int_region_bytes(int nbatch) {
  return align_up(sizeof(int) * (HEADER_INTS + 2 * (size_t)nbatch), 256);
}

// workspace = [status | queue | class_range | hist | cursor | nreal | order | scratch]
Work carve_work(void* ws, int nbatch) {
  Work wk;
  int* p = reinterpret_cast<int*>(ws);
  wk.status = p, p += 1;
  wk.done = p, p += 1;
  wk.queue = p, p += NCLASS + 1;
  wk.class_range = p, p += 2 * (NCLASS + 1);
  wk.hist = p, p += HIST_BINS;
  wk.cursor = p, p += HIST_BINS;
  wk.nreal = p, p += nbatch;
  wk.order = p, p += nbatch;
  return wk;
}

constexpr int MAX_SMS = 160;

size_t scratch_bytes_class(int c, int cap, size_t elem) {
  // per CTA: five pair planes + the D4S weight table [cap][D4S_ECAP][D4S_WSTR] (Lay<>::scratch_stride)
  const size_t per_cta = (size_t)5 * (cap * (cap - 1) / 2) + (size_t)cap * D4S_ECAP * D4S_WSTR;
  return align_up((size_t)MAX_SMS * class_occ_cap(c) * per_cta * elem, 256);
}

template <typename T>
Par<T> make_par(const d4b200_params* p, double ga) {
  Par<T> P;
  P.s6 = (T)p->s6;
  P.s8 = (T)p->s8;
  P.s10k = p->has_s10 ? (T)(p->s10 * 49.0 / 40.0) : T(0);
  P.a1 = (T)p->a1;
  P.a2 = (T)p->a2;
  P.alp3 = (T)(p->alp / 3.0);
  P.fac9 = (T)cbrt(p->s9 / 6.0);
  P.disp2_sq = (T)(p->disp2_cutoff * p->disp2_cutoff);
  P.disp3_sq = (T)(p->disp3_cutoff * p->disp3_cutoff);
  P.cn_sq = (T)(p->cn_cutoff * p->cn_cutoff);
  P.wf = p->wf;
  P.ga = ga;
  P.has_atm = p->s9 != 0.0;
  P.alp16 = p->alp == 16.0;
  P.model = p->model;
  return P;
}

template <typename T, bool GRAD>
int configure(d4b200_tables* h, int model) {
  return flavour_configure<T>(h, GRAD, model);
}

template <typename T, bool GRAD>
int run_small(d4b200_tables* h, const d4b200_params* par, int nbatch, int nat,
              const int64_t* numbers, const T* pos, const T* q, const T* gin, T* energy, T* cn_out,
              T* grad, T* gradq, void* ws, size_t ws_bytes, cudaStream_t st, T* c6_out = nullptr,
              T* alpha_out = nullptr, int pool = 0) {
  constexpr int dt = sizeof(T) == 8 ? 0 : 1;
  constexpr int gr = GRAD ? 1 : 0;
  g_launches = 0;
  const int md = par && par->model == D4B200_MODEL_D4S ? 1 : 0;
  if (!h || !par || !numbers || !pos || !q || !ws || nbatch < 0 || nat < 0) return D4B200_EINVAL;
  if (!GRAD && !energy) return D4B200_EINVAL;
  if (!(par->a1 == par->a1) || !(par->a2 == par->a2)) return D4B200_EPARAM;
  if (ws_bytes < d4b200_workspace_bytes(nbatch, nat)) return D4B200_EWORKSPACE;
  if (nbatch == 0 || nat == 0) return 0;

  Work wk = carve_work(ws, nbatch);
  if (h->profile) cudaEventRecord(h->ev_call[0], st);
  cudaError_t e = cudaMemsetAsync(wk.status, 0, sizeof(int) * HEADER_INTS, st);
  if (e != cudaSuccess) return (int)e;
  ++g_launches;
  Caps caps;
  for (int c = 0; c < NCLASS; ++c) caps.v[c] = h->caps[md][dt][gr][c];
  if (nbatch <= PLAN_MAX_BATCH) {
    k_plan<<<(nbatch * 32 + PLAN_THREADS - 1) / PLAN_THREADS, PLAN_THREADS, 0, st>>>(numbers, nbatch, nat, wk, caps);
    g_launches += 1;
  } else {
    k_count<<<(nbatch + 7) / 8, 256, 0, st>>>(numbers, nbatch, nat, wk);
    k_scan<<<1, 32, 0, st>>>(nbatch, wk, caps);
    k_scatter<<<(nbatch + 255) / 256, 256, 0, st>>>(nbatch, wk);
    g_launches += 3;
  }

  for (int c = 0; c < NCLASS; ++c) h->ev_used[c] = 0;
  SmallArgs<T> A;
  A.numbers = numbers;
  A.pos = pos;
  A.q = q;
  A.gin = gin;
  A.energy = energy;
  A.cn_out = cn_out;
  A.grad = grad;
  A.gradq = gradq;
  A.c6_out = c6_out;
  A.alpha_out = alpha_out;
  A.nbatch = nbatch;
  A.phase = nullptr;
  A.nat = nat;
  if constexpr (dt == 0) {
    A.tab = h->t64;
  } else {
    A.tab = h->t32;
  }
  A.par = make_par<T>(par, h->ga);
  A.wk = wk;
  unsigned char* scratch = reinterpret_cast<unsigned char*>(ws) + int_region_bytes(nbatch);
  // fork: every populated class runs on its own stream, ordered after the prep kernels
  if (h->profile) cudaEventRecord(h->ev_call[1], st);
  cudaEventRecord(h->ev_fork[pool], st);
  // (measured: launching the small classes first, all streams at equal priority, is 2-3 %
  // faster on C2 than large-first with prioritised streams)
  size_t soff[NCLASS];
  int lows[NCLASS];
  {
    int prev = 0;
    size_t off = 0;
    for (int c = 0; c < NCLASS; ++c) {
      lows[c] = prev + 1;
      prev = caps.v[c];
      soff[c] = off;
      off += scratch_bytes_class(c, caps.v[c], sizeof(T));
    }
  }
  for (int c = 0; c < NCLASS; ++c) {
    // a class can only be populated if the padded width reaches into it
    if (nat >= lows[c] || c == 0) {
      A.cls = c;
      A.phase = h->phase_on ? h->phase_dev + 16 * c : nullptr;
      A.scratch = reinterpret_cast<T*>(scratch + soff[c]);
      long grid = (long)h->grid_per_sm[md][dt][gr][c] * h->num_sms;
      if (grid > nbatch) grid = nbatch;
      const long gmax = (long)(h->num_sms < MAX_SMS ? h->num_sms : MAX_SMS) * class_occ_cap(c);
      if (grid > gmax) grid = gmax;
      cudaStream_t cs = h->profile ? st : h->cstream[pool][c];  // profiling: serialise on the caller's stream
      if (!h->profile) cudaStreamWaitEvent(cs, h->ev_fork[pool], 0);
      if (h->profile) cudaEventRecord(h->ev[2 * c], st);
      flavour_launch<T>(GRAD, par->model, c, (unsigned)grid, cs, A);
      if (h->profile) cudaEventRecord(h->ev[2 * c + 1], st);
      h->ev_used[c] = h->profile;
      if (!h->profile) {
        cudaEventRecord(h->ev_join[pool][c], cs);
        cudaStreamWaitEvent(st, h->ev_join[pool][c], 0);
      }
      ++g_launches;
    }
  }
  if (h->profile) cudaEventRecord(h->ev_call[2], st);
  g_total_launches += g_launches;
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace

// Host-buffer variant of the energy call: the batch is cut into chunks that flow through
// D4_HOST_SLOTS pipeline slots (H2D copy -> kernels -> D2H copy), so the copies of one chunk
// overlap the kernels of the others.  Pinned host memory is needed for the overlap (pageable
// memory still works, serialised by the driver).  Synchronous: returns when
// ``energy_host`` is complete.
// GRAD: the fused energy + gradient kernels (d(sum E)/d positions, optionally d(sum E)/dq) run
// per chunk and the gradient planes travel back with the energies.
// atomic numbers that travelled as 1- or 4-byte integers: widen on the device (behind the copy)
template <typename Z>
__global__ void k_widen(const Z* __restrict__ src, int64_t* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = (int64_t)src[i];
}

// ``zsize``: bytes per atomic number in the caller's host array (8 = int64, 4 = int32, 1 = uint8).
// Narrow arrays are uploaded as they are -- one eighth of the bytes for uint8, no host-side pass --
// and widened by a small kernel on the chunk's stream.  ``status_out`` (optional) receives the OR of
// the kernels' status words of all chunks (D4B200_STATUS_*).
template <typename T, bool GRAD>
static int run_energy_host(d4b200_tables* h, const d4b200_params* par, int nbatch, int nat,
                           const void* numbers_any, int zsize, const T* pos, const T* q, T* energy, int chunks,
                           T* grad = nullptr, T* gradq = nullptr, int* status_out = nullptr) {
  const unsigned char* numbers = reinterpret_cast<const unsigned char*>(numbers_any);
  if (!h || !par || !numbers || !pos || !q || !energy || nbatch < 0 || nat < 0) return D4B200_EINVAL;
  if (zsize != 8 && zsize != 4 && zsize != 1) return D4B200_EINVAL;
  if (GRAD && !grad) return D4B200_EINVAL;
  if (nbatch == 0 || nat == 0) return 0;
  int prev_dev = 0;
  cudaGetDevice(&prev_dev);
  cudaSetDevice(h->device);
  if (chunks <= 0) chunks = nbatch >= 2048 ? 4 : nbatch >= 512 ? (GRAD ? 4 : 2) : 1;
  if (chunks > nbatch) chunks = nbatch;
  if (chunks > D4_HOST_STATUS) chunks = D4_HOST_STATUS;
  const int cb = (nbatch + chunks - 1) / chunks;  // structures per chunk
  const size_t rows = (size_t)cb * nat;
  const size_t off_pos = align_up(rows * sizeof(int64_t), 256);
  const size_t off_q = off_pos + align_up(rows * 3 * sizeof(T), 256);
  const size_t off_e = off_q + align_up(rows * sizeof(T), 256);
  const size_t off_g = off_e + align_up(rows * sizeof(T), 256);
  const size_t off_gq = off_g + (GRAD ? align_up(rows * 3 * sizeof(T), 256) : 0);
  const size_t off_zn = off_gq + (GRAD ? align_up(rows * sizeof(T), 256) : 0);  // narrow numbers as uploaded
  const size_t off_ws = off_zn + (zsize != 8 ? align_up(rows * zsize, 256) : 0);
  const size_t ws_bytes = d4b200_workspace_bytes(cb, nat);
  const size_t need = off_ws + ws_bytes;
  cudaError_t e = cudaSuccess;
  if (!h->hstatus) e = cudaHostAlloc(&h->hstatus, sizeof(int) * D4_HOST_STATUS, cudaHostAllocDefault);

  if (!h->hcopy) e = cudaStreamCreateWithFlags(&h->hcopy, cudaStreamNonBlocking);
  for (int s = 0; s < D4_HOST_SLOTS && s < chunks && e == cudaSuccess; ++s) {
    if (!h->hstream[s]) {
      e = cudaStreamCreateWithFlags(&h->hstream[s], cudaStreamNonBlocking);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->hev_in[s], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->hev_done[s], cudaEventDisableTiming);
    }
    if (e == cudaSuccess && h->hbuf_bytes[s] < need) {
      if (h->hbuf[s]) cudaFree(h->hbuf[s]);
      h->hbuf[s] = nullptr;
      h->hbuf_bytes[s] = 0;
      e = cudaMalloc(&h->hbuf[s], need);
      if (e == cudaSuccess) h->hbuf_bytes[s] = need;
    }
  }
  int rc = 0;
  long long launches = 0;
  for (int c = 0; c < chunks && e == cudaSuccess && rc == 0; ++c) {
    const int b0 = c * cb, nb = (b0 + cb <= nbatch ? cb : nbatch - b0);
    if (nb <= 0) break;
    const int s = c % D4_HOST_SLOTS;
    unsigned char* d = reinterpret_cast<unsigned char*>(h->hbuf[s]);
    cudaStream_t st = h->hstream[s];
    const size_t r = (size_t)nb * nat, o = (size_t)b0 * nat;
    // all uploads share one stream so that they reach the device in chunk order at full
    // link bandwidth; the slot's stream picks up when its chunk has landed
    if (c >= D4_HOST_SLOTS) cudaStreamWaitEvent(h->hcopy, h->hev_done[s], 0);
    cudaMemcpyAsync(zsize == 8 ? d : d + off_zn, numbers + o * zsize, r * zsize, cudaMemcpyHostToDevice, h->hcopy);
    cudaMemcpyAsync(d + off_pos, pos + 3 * o, r * 3 * sizeof(T), cudaMemcpyHostToDevice, h->hcopy);
    cudaMemcpyAsync(d + off_q, q + o, r * sizeof(T), cudaMemcpyHostToDevice, h->hcopy);
    cudaEventRecord(h->hev_in[s], h->hcopy);
    cudaStreamWaitEvent(st, h->hev_in[s], 0);
    if (zsize != 8) {
      const unsigned wg = (unsigned)((r + 255) / 256 < 1184 ? (r + 255) / 256 : 1184);
      if (zsize == 4)
        k_widen<int32_t><<<wg, 256, 0, st>>>(reinterpret_cast<const int32_t*>(d + off_zn), reinterpret_cast<int64_t*>(d), r);
      else
        k_widen<uint8_t><<<wg, 256, 0, st>>>(d + off_zn, reinterpret_cast<int64_t*>(d), r);
    }
    rc = run_small<T, GRAD>(h, par, nb, nat, reinterpret_cast<const int64_t*>(d),
                            reinterpret_cast<const T*>(d + off_pos), reinterpret_cast<const T*>(d + off_q),
                            nullptr, reinterpret_cast<T*>(d + off_e), nullptr,
                            GRAD ? reinterpret_cast<T*>(d + off_g) : nullptr,
                            GRAD && gradq ? reinterpret_cast<T*>(d + off_gq) : nullptr, d + off_ws, ws_bytes,
                            st, nullptr, nullptr, 1 + s);
    launches += g_launches + (zsize != 8 ? 1 : 0);
    // status word of the chunk (first int of its workspace, see d4b200_status) -> pinned host slot
    cudaMemcpyAsync(h->hstatus + c, d + off_ws, sizeof(int), cudaMemcpyDeviceToHost, st);
    e = cudaMemcpyAsync(energy + o, d + off_e, r * sizeof(T), cudaMemcpyDeviceToHost, st);
    if (GRAD) {
      cudaMemcpyAsync(grad + 3 * o, d + off_g, r * 3 * sizeof(T), cudaMemcpyDeviceToHost, st);
      if (gradq) cudaMemcpyAsync(gradq + o, d + off_gq, r * sizeof(T), cudaMemcpyDeviceToHost, st);
    }
    cudaEventRecord(h->hev_done[s], st);
  }
  for (int s = 0; s < D4_HOST_SLOTS; ++s)
    if (h->hstream[s]) {
      cudaError_t es = cudaStreamSynchronize(h->hstream[s]);
      if (e == cudaSuccess) e = es;
    }
  g_launches = (int)launches;
  g_total_launches += (zsize != 8 ? chunks : 0);
  int bits = 0;
  if (rc == 0 && e == cudaSuccess)
    for (int c = 0; c < chunks; ++c) bits |= h->hstatus[c];
  if (status_out) *status_out = bits;
  cudaSetDevice(prev_dev);
  if (rc != 0) return rc;
  if (e != cudaSuccess) return (int)e;
  // an atomic number outside 1..103 / a structure beyond the kernels' limit is an error of the call
  // (the device entry points report it through d4b200_status)
  if (bits & D4B200_STATUS_BAD_NUMBER) return D4B200_ENUMBER;
  if (bits & D4B200_STATUS_TOO_LARGE) return D4B200_ETOOLARGE;
  return 0;
}

// (cn, C6, alpha) of tad_dftd4.get_properties (disp.py:149-197); the energy kernel stops
// after the weighted-polarizability vectors.  ``energy_scratch_dev`` [nbatch, nat] is zeroed.
template <typename T>
static int run_props(d4b200_tables_t t, const d4b200_params* par, int nbatch, int nat,
                     const int64_t* numbers, const T* pos, const T* q, T* cn, T* c6, T* alpha,
                     T* energy_scratch, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!c6 || !energy_scratch) return D4B200_EINVAL;
  cudaError_t e = cudaMemsetAsync(c6, 0, sizeof(T) * (size_t)nbatch * nat * nat, st);
  if (e != cudaSuccess) return (int)e;
  return run_small<T, false>(t, par, nbatch, nat, numbers, pos, q, nullptr, energy_scratch, cn, nullptr,
                             nullptr, ws, ws_bytes, st, c6, alpha);
}

extern "C" {

int d4b200_version(void) { return D4B200_VERSION; }

const char* d4b200_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case D4B200_EINVAL: return "invalid argument (null pointer or negative size)";
    case D4B200_EWORKSPACE: return "workspace too small";
    case D4B200_EPARAM: return "damping parameters a1/a2 missing";
    case D4B200_ETABLE: return "table blob has the wrong size";
    case D4B200_EARCH: return "device is not sm_100 (B200)";
    case D4B200_ENUMBER: return "numbers contains an atomic number outside 1..103 (0 = padding)";
    case D4B200_ETOOLARGE: return "a structure is larger than the one-CTA-per-structure kernels support";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

int d4b200_tables_create(int device, const double* f64_blob_host, size_t n_f64,
                         const int32_t* i32_blob_host, size_t n_i32, double ga, double gc,
                         d4b200_tables_t* out) {
  if (!f64_blob_host || !i32_blob_host || !out) return D4B200_EINVAL;
  const BlobOffsets o = blob_offsets();
  if (n_f64 != o.total || n_i32 != o.itotal) return D4B200_ETABLE;
  int prev_dev = 0;
  cudaError_t e = cudaGetDevice(&prev_dev);
  if (e != cudaSuccess) return (int)e;
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return (int)e;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return (int)e;
  if (prop.major != 10) {
    cudaSetDevice(prev_dev);
    return D4B200_EARCH;
  }
  d4b200_tables* h = new (std::nothrow) d4b200_tables();
  if (!h) return D4B200_EINVAL;
  memset(h, 0, sizeof(*h));
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->ga = ga;
  h->gc = gc;
  h->n_f64 = n_f64;
  h->n_i32 = n_i32;
  int rc = 0;
  do {
    if ((e = cudaMalloc(&h->f64, n_f64 * sizeof(double))) != cudaSuccess) break;
    if ((e = cudaMalloc(&h->f32, n_f64 * sizeof(float))) != cudaSuccess) break;
    if ((e = cudaMalloc(&h->i32, n_i32 * sizeof(int))) != cudaSuccess) break;
    if ((e = cudaMemcpy(h->f64, f64_blob_host, n_f64 * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) break;
    if ((e = cudaMemcpy(h->i32, i32_blob_host, n_i32 * sizeof(int), cudaMemcpyHostToDevice)) != cudaSuccess) break;
    if ((e = cudaMalloc(&h->pij, sizeof(unsigned short) * SMALL_MAX * (SMALL_MAX - 1) / 2)) != cudaSuccess) break;
    k_pair_table<<<SMALL_MAX - 1, 128>>>(h->pij);
    k_to_float<<<(unsigned)((n_f64 + 255) / 256), 256>>>(h->f64, h->f32, n_f64);
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) break;
    for (int p = 0; p < 1 + D4_HOST_SLOTS && e == cudaSuccess; ++p) {
      for (int c = 0; c < NCLASS && e == cudaSuccess; ++c) {
        e = cudaStreamCreateWithFlags(&h->cstream[p][c], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join[p][c], cudaEventDisableTiming);
      }
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork[p], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) break;
    h->t64 = make_tables<double>(h->f64, h->f64, h->i32, h->pij);
    h->t32 = make_tables<float>(h->f32, h->f64, h->i32, h->pij);
    for (int model = 0; model < 2 && rc == 0; ++model) {
      if ((rc = configure<double, false>(h, model)) != 0) break;
      if ((rc = configure<double, true>(h, model)) != 0) break;
      if ((rc = configure<float, false>(h, model)) != 0) break;
      if ((rc = configure<float, true>(h, model)) != 0) break;
    }
    if (rc != 0) break;
  } while (0);
  cudaSetDevice(prev_dev);
  if (e != cudaSuccess || rc != 0) {
    d4b200_tables_destroy(h);
    return e != cudaSuccess ? (int)e : rc;
  }
  *out = h;
  return 0;
}

int d4b200_tables_destroy(d4b200_tables_t h) {
  if (!h) return 0;
  for (int i = 0; i < 2 * NCLASS; ++i)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  for (int i = 0; i < 3; ++i)
    if (h->ev_call[i]) cudaEventDestroy(h->ev_call[i]);
  for (int p = 0; p < 1 + D4_HOST_SLOTS; ++p) {
    for (int c = 0; c < NCLASS; ++c) {
      if (h->cstream[p][c]) cudaStreamDestroy(h->cstream[p][c]);
      if (h->ev_join[p][c]) cudaEventDestroy(h->ev_join[p][c]);
    }
    if (h->ev_fork[p]) cudaEventDestroy(h->ev_fork[p]);
  }
  cudaFree(h->phase_dev);
  cudaFree(h->pij);
  if (h->hstatus) cudaFreeHost(h->hstatus);
  if (h->hcopy) cudaStreamDestroy(h->hcopy);
  for (int s = 0; s < D4_HOST_SLOTS; ++s) {
    if (h->hbuf[s]) cudaFree(h->hbuf[s]);
    if (h->hstream[s]) {
      cudaStreamDestroy(h->hstream[s]);
      cudaEventDestroy(h->hev_in[s]);
      cudaEventDestroy(h->hev_done[s]);
    }
  }
  cudaFree(h->f64);
  cudaFree(h->f32);
  cudaFree(h->i32);
  delete h;
  return 0;
}

size_t d4b200_workspace_bytes(int nbatch, int nat) {
  (void)nat;
  if (nbatch < 0) return 0;
  size_t s = int_region_bytes(nbatch);
  const int caps[NCLASS] = {32, 48, 64, 96, 128};
  for (int c = 0; c < NCLASS; ++c) s += scratch_bytes_class(c, caps[c], sizeof(double));
  return s;
}

int d4b200_energy_f64(d4b200_tables_t t, const d4b200_params* par, int nbatch, int nat,
                      const int64_t* numbers, const double* pos, const double* q, double* energy,
                      double* cn, void* ws, size_t ws_bytes, void* stream) {
  return run_small<double, false>(t, par, nbatch, nat, numbers, pos, q, nullptr, energy, cn,
                                  nullptr, nullptr, ws, ws_bytes, (cudaStream_t)stream);
}
int d4b200_energy_f32(d4b200_tables_t t, const d4b200_params* par, int nbatch, int nat,
                      const int64_t* numbers, const float* pos, const float* q, float* energy,
                      float* cn, void* ws, size_t ws_bytes, void* stream) {
  return run_small<float, false>(t, par, nbatch, nat, numbers, pos, q, nullptr, energy, cn, nullptr,
                                 nullptr, ws, ws_bytes, (cudaStream_t)stream);
}
int d4b200_energy_host_f64(d4b200_tables_t t, const d4b200_params* par, int nbatch, int nat,
                           const int64_t* numbers_host, const double* pos_host,
                           const double* q_host, double* energy_host, int chunks) {
  return run_energy_host<double, false>(t, par, nbatch, nat, numbers_host, 8, pos_host, q_host, energy_host, chunks);
}
int d4b200_energy_host_f32(d4b200_tables_t t, const d4b200_params* par, int nbatch, int nat,
                           const int64_t* numbers_host, const float* pos_host, const float* q_host,
                           float* energy_host, int chunks) {
  return run_energy_host<float, false>(t, par, nbatch, nat, numbers_host, 8, pos_host, q_host, energy_host, chunks);
}
int d4b200_energy_gradient_host_f64(d4b200_tables_t t, const d4b200_params* par, int nbatch, int nat,
                                    const int64_t* numbers_host, const double* pos_host,
                                    const double* q_host, double* energy_host, double* grad_host,
                                    double* gradq_host, int chunks) {
  return run_energy_host<double, true>(t, par, nbatch, nat, numbers_host, 8, pos_host, q_host, energy_host, chunks,
                                       grad_host, gradq_host);
}
int d4b200_energy_gradient_host_f32(d4b200_tables_t t, const d4b200_params* par, int nbatch, int nat,
                                    const int64_t* numbers_host, const float* pos_host, const float* q_host,
                                    float* energy_host, float* grad_host, float* gradq_host, int chunks) {
  return run_energy_host<float, true>(t, par, nbatch, nat, numbers_host, 8, pos_host, q_host, energy_host, chunks,
                                      grad_host, gradq_host);
}
int d4b200_energy_host_z_f64(d4b200_tables_t t, const d4b200_params* par, int nbatch, int nat,
                             const void* numbers_host, int numbers_itemsize, const double* pos_host,
                             const double* q_host, double* energy_host, double* grad_host, double* gradq_host,
                             int chunks, int* status_out) {
  if (grad_host)
    return run_energy_host<double, true>(t, par, nbatch, nat, numbers_host, numbers_itemsize, pos_host, q_host,
                                         energy_host, chunks, grad_host, gradq_host, status_out);
  return run_energy_host<double, false>(t, par, nbatch, nat, numbers_host, numbers_itemsize, pos_host, q_host,
                                        energy_host, chunks, nullptr, nullptr, status_out);
}
int d4b200_energy_host_z_f32(d4b200_tables_t t, const d4b200_params* par, int nbatch, int nat,
                             const void* numbers_host, int numbers_itemsize, const float* pos_host,
                             const float* q_host, float* energy_host, float* grad_host, float* gradq_host,
                             int chunks, int* status_out) {
  if (grad_host)
    return run_energy_host<float, true>(t, par, nbatch, nat, numbers_host, numbers_itemsize, pos_host, q_host,
                                        energy_host, chunks, grad_host, gradq_host, status_out);
  return run_energy_host<float, false>(t, par, nbatch, nat, numbers_host, numbers_itemsize, pos_host, q_host,
                                       energy_host, chunks, nullptr, nullptr, status_out);
}
int d4b200_gradient_f64(d4b200_tables_t t, const d4b200_params* par, int nbatch, int nat,
                        const int64_t* numbers, const double* pos, const double* q,
                        const double* gin, double* grad, double* gradq, void* ws, size_t ws_bytes,
                        void* stream) {
  return run_small<double, true>(t, par, nbatch, nat, numbers, pos, q, gin, nullptr, nullptr, grad,
                                 gradq, ws, ws_bytes, (cudaStream_t)stream);
}
int d4b200_gradient_f32(d4b200_tables_t t, const d4b200_params* par, int nbatch, int nat,
                        const int64_t* numbers, const float* pos, const float* q, const float* gin,
                        float* grad, float* gradq, void* ws, size_t ws_bytes, void* stream) {
  return run_small<float, true>(t, par, nbatch, nat, numbers, pos, q, gin, nullptr, nullptr, grad,
                                gradq, ws, ws_bytes, (cudaStream_t)stream);
}
int d4b200_energy_gradient_f64(d4b200_tables_t t, const d4b200_params* par, int nbatch, int nat,
                               const int64_t* numbers, const double* pos, const double* q,
                               const double* gin, double* energy, double* grad, double* gradq,
                               void* ws, size_t ws_bytes, void* stream) {
  if (!energy) return D4B200_EINVAL;
  return run_small<double, true>(t, par, nbatch, nat, numbers, pos, q, gin, energy, nullptr, grad,
                                 gradq, ws, ws_bytes, (cudaStream_t)stream);
}
int d4b200_energy_gradient_f32(d4b200_tables_t t, const d4b200_params* par, int nbatch, int nat,
                               const int64_t* numbers, const float* pos, const float* q,
                               const float* gin, float* energy, float* grad, float* gradq, void* ws,
                               size_t ws_bytes, void* stream) {
  if (!energy) return D4B200_EINVAL;
  return run_small<float, true>(t, par, nbatch, nat, numbers, pos, q, gin, energy, nullptr, grad,
                                gradq, ws, ws_bytes, (cudaStream_t)stream);
}

int d4b200_properties_f64(d4b200_tables_t t, const d4b200_params* par, int nbatch, int nat,
                          const int64_t* numbers, const double* pos, const double* q, double* cn,
                          double* c6, double* alpha, double* energy_scratch, void* ws,
                          size_t ws_bytes, void* stream) {
  return run_props<double>(t, par, nbatch, nat, numbers, pos, q, cn, c6, alpha, energy_scratch, ws,
                           ws_bytes, (cudaStream_t)stream);
}
int d4b200_properties_f32(d4b200_tables_t t, const d4b200_params* par, int nbatch, int nat,
                          const int64_t* numbers, const float* pos, const float* q, float* cn,
                          float* c6, float* alpha, float* energy_scratch, void* ws, size_t ws_bytes,
                          void* stream) {
  return run_props<float>(t, par, nbatch, nat, numbers, pos, q, cn, c6, alpha, energy_scratch, ws,
                          ws_bytes, (cudaStream_t)stream);
}

int d4b200_status(void* ws, void* stream, int* bits) {
  if (!ws || !bits) return D4B200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemcpyAsync(bits, ws, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaStreamSynchronize(st);
  return e == cudaSuccess ? 0 : (int)e;
}

int d4b200_last_launch_count(void) { return g_launches; }
long long d4b200_total_launch_count(void) { return g_total_launches; }

// development profiling: per-phase cycle counters of the small-family kernels
int d4b200_phase_profile(d4b200_tables_t h, int enable, unsigned long long* out /*[NCLASS*16] or NULL*/) {
  if (!h) return D4B200_EINVAL;
  if (!h->phase_dev) {
    cudaError_t e = cudaMalloc(&h->phase_dev, sizeof(unsigned long long) * 16 * NCLASS);
    if (e != cudaSuccess) return (int)e;
    cudaMemset(h->phase_dev, 0, sizeof(unsigned long long) * 16 * NCLASS);
  }
  if (out) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return (int)e;
    cudaMemcpy(out, h->phase_dev, sizeof(unsigned long long) * 16 * NCLASS, cudaMemcpyDeviceToHost);
    cudaMemset(h->phase_dev, 0, sizeof(unsigned long long) * 16 * NCLASS);
  }
  h->phase_on = enable != 0;
  return 0;
}

int d4b200_profile_enable(d4b200_tables_t h, int enable) {
  if (!h) return D4B200_EINVAL;
  if (enable) {
    for (int i = 0; i < 2 * NCLASS; ++i) {
      if (!h->ev[i]) {
        cudaError_t e = cudaEventCreate(&h->ev[i]);
        if (e != cudaSuccess) return (int)e;
      }
    }
    for (int i = 0; i < 3; ++i) {
      if (!h->ev_call[i]) {
        cudaError_t e = cudaEventCreate(&h->ev_call[i]);
        if (e != cudaSuccess) return (int)e;
      }
    }
  }
  h->profile = enable != 0;
  return 0;
}

int d4b200_profile_read(d4b200_tables_t h, float* ms_per_class) {
  if (!h || !ms_per_class) return D4B200_EINVAL;
  for (int c = 0; c < NCLASS; ++c) {
    ms_per_class[c] = -1.0f;
    if (!h->ev_used[c]) continue;
    cudaError_t e = cudaEventSynchronize(h->ev[2 * c + 1]);
    if (e != cudaSuccess) return (int)e;
    e = cudaEventElapsedTime(&ms_per_class[c], h->ev[2 * c], h->ev[2 * c + 1]);
    if (e != cudaSuccess) return (int)e;
  }
  cudaError_t e = cudaEventSynchronize(h->ev_call[2]);
  if (e != cudaSuccess) return (int)e;
  cudaEventElapsedTime(&ms_per_class[NCLASS], h->ev_call[0], h->ev_call[1]);
  cudaEventElapsedTime(&ms_per_class[NCLASS + 1], h->ev_call[0], h->ev_call[2]);
  return 0;
}

int d4b200_class_caps(d4b200_tables_t h, int fp32, int grad, int* caps_out) {
  if (!h || !caps_out) return D4B200_EINVAL;
  for (int c = 0; c < NCLASS; ++c) caps_out[c] = h->caps[0][fp32 ? 1 : 0][grad ? 1 : 0][c];
  return 0;
}

int d4b200_class_caps_model(d4b200_tables_t h, int fp32, int grad, int model, int* caps_out) {
  if (!h || !caps_out || model < 0 || model > 1) return D4B200_EINVAL;
  for (int c = 0; c < NCLASS; ++c) caps_out[c] = h->caps[model][fp32 ? 1 : 0][grad ? 1 : 0][c];
  return 0;
}

int d4b200_small_limit(d4b200_tables_t h, int fp32, int grad, int model) {
  if (!h || model < 0 || model > 1) return D4B200_EINVAL;
  return h->caps[model][fp32 ? 1 : 0][grad ? 1 : 0][NCLASS - 1];
}

// FP64 vector (DFMA) throughput of the device, measured: the roofline
// denominator for the FP64-bound kernels (MEASURED_PEAKS.json has no FP64 entry).
__global__ void k_dfma_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c), a1 = fma(a1, m, c), a2 = fma(a2, m, c), a3 = fma(a3, m, c);
    a4 = fma(a4, m, c), a5 = fma(a5, m, c), a6 = fma(a6, m, c), a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int d4b200_measure_fp64_peak(d4b200_tables_t h, void* scratch_dev, size_t scratch_bytes,
                             void* stream, double* tflops_out) {
  if (!h || !scratch_dev || !tflops_out) return D4B200_EINVAL;
  const int blocks = h->num_sms * 8, threads = 256, iters = 1 << 15;
  if (scratch_bytes < (size_t)blocks * threads * sizeof(double)) return D4B200_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, st);
    k_dfma_peak<<<blocks, threads, 0, st>>>((double*)scratch_dev, iters);
    cudaEventRecord(e1, st);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) return (int)e;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops_out = best;
  return 0;
}

}  // extern "C"

// Model-level entry points: the reference's D4Model / D4SModel methods as stand-alone device
// kernels (the fused energy/gradient kernels evaluate the same quantities internally and
// never materialise them):
//   weight_references   (src/tad_dftd4/model/d4.py:103-228, model/d4s.py:109-266)
//   get_atomic_c6       (model/d4.py:268-289, model/d4s.py:268-290)
//   get_weighted_pols / get_polarizabilities (model/d4.py:291-307, model/base.py:286-302)
// Gather/contract kernels, one thread per output element; HBM-bound (the (N,N,7) D4S
// weights and the (N,N) C6 matrix are the traffic).
#include <cuda_runtime.h>
#include <math.h>

#include "d4b200_handle.cuh"
#include "d4b200_small.cuh"

namespace {

using namespace d4b200;

// zeta(gam gc, refq + zeff, q + zeff) of model/base.py:326-335 and its q-derivative
__device__ __forceinline__ void zeta_of(double ga, double gam, double qref, double qmod, double eps, bool on,
                                        double& zeta, double& dzeta) {
  const bool qpos = qmod > 0.0;
  const double qinv = 1.0 / (qpos ? qmod - eps : 1.0);
  const double scale = qpos ? exp(gam * (1.0 - qref * qinv)) : 0.0;
  zeta = on ? exp(ga * (1.0 - scale)) : 0.0;
  dzeta = -ga * gam * scale * zeta * qref * qinv * qinv;
}

// D4: thread per (structure, atom); D4S: thread per (structure, partner m, atom n)
template <typename T, bool D4S>
__global__ void k_weight_references(Tables<T> tab, double ga, double wf, long long total, int nat,
                                    const int64_t* __restrict__ numbers, const T* __restrict__ cn,
                                    const T* __restrict__ q, T* __restrict__ gw, T* __restrict__ dgwdcn,
                                    T* __restrict__ dgwdq, int* __restrict__ status) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  long long atom = t;  // flat index of the weighted atom n in [nbatch, nat]
  double w = wf;
  long long zn = 0;
  if (D4S) {
    const long long n = t % nat, bm = t / nat;
    atom = (bm / nat) * nat + n;
    zn = numbers[atom];
    long long zm = numbers[bm];
    if (zm < 0 || zm >= NELEM) zm = 0;
    if (zn >= 0 && zn < NELEM) w = tab.wfpair[zn * NELEM + zm];
  } else {
    zn = numbers[atom];
  }
  if (zn < 0 || zn >= NELEM) {
    if (status) atomicOr(status, D4B200_STATUS_BAD_NUMBER);
    zn = 0;
  }
  const int z = (int)zn;
  const double cni = cn ? (double)cn[atom] : 0.0;
  const double qi = q ? (double)q[atom] : 0.0;
  double S[NREF], dS[NREF], norm, dnorm;
  d4s_gauss<true>(tab.refcn, tab.refc, z, cni, w, S, dS, norm, dnorm);
  const double inv = norm > 0.0 ? 1.0 / norm : 0.0;
  const double gam = tab.gamgc[z], zeff = tab.zeff[z];
#pragma unroll
  for (int a = 0; a < NREF; ++a) {
    const bool on = tab.refc[z * NREF + a] > 0;
    double zeta, dzeta;
    zeta_of(ga, gam, tab.refq[z * NREF + a], qi + zeff, (double)d4_eps<T>(), on, zeta, dzeta);
    const double g = S[a] * inv;
    const double dg = (dS[a] - g * dnorm) * inv;
    gw[t * NREF + a] = (T)(zeta * g);
    if (dgwdcn) dgwdcn[t * NREF + a] = (T)(zeta * dg);
    if (dgwdq) dgwdq[t * NREF + a] = (T)(on ? dzeta * g : 0.0);
  }
}

// thread per (structure, i, j)
template <typename T, bool D4S>
__global__ void k_atomic_c6(Tables<T> tab, long long total, int nat, const int64_t* __restrict__ numbers,
                            const T* __restrict__ gw, T* __restrict__ c6) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const long long j = t % nat, bi = t / nat, b = bi / nat, i = bi % nat;
  long long zi = numbers[bi], zj = numbers[b * nat + j];
  if (zi < 0 || zi >= NELEM) zi = 0;
  if (zj < 0 || zj >= NELEM) zj = 0;
  const T* R = tab.rc6 + ((size_t)zi * NELEM + zj) * (NREF * NREF);
  // D4: gw[b, i, :], gw[b, j, :];  D4S: gw[b, j, i, :] (atom i seen by j), gw[b, i, j, :]
  const T* wi = D4S ? gw + ((b * nat + j) * nat + i) * NREF : gw + bi * NREF;
  const T* wj = D4S ? gw + ((b * nat + i) * nat + j) * NREF : gw + (b * nat + j) * NREF;
  T vj[NREF];
#pragma unroll
  for (int bq = 0; bq < NREF; ++bq) vj[bq] = wj[bq];
  T acc = T(0);
#pragma unroll
  for (int a = 0; a < NREF; ++a) {
    T s = T(0);
#pragma unroll
    for (int bq = 0; bq < NREF; ++bq) s += R[a * NREF + bq] * vj[bq];
    acc += wi[a] * s;
  }
  c6[t] = acc;
}

struct FreqNorm {
  double inv[NFREQ];  // 1 / sqrt(3/pi * trapezoid weight): alpha = alpha_w * inv
};

// thread per (structure, atom, frequency)
template <typename T>
__global__ void k_weighted_pols(Tables<T> tab, FreqNorm fn, long long total, int nfreq,
                                const int64_t* __restrict__ numbers, const T* __restrict__ gw,
                                T* __restrict__ alpha) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int w = (int)(t % nfreq);
  const long long atom = t / nfreq;
  long long z = numbers[atom];
  if (z < 0 || z >= NELEM) z = 0;
  T acc = T(0);
#pragma unroll
  for (int a = 0; a < NREF; ++a) {
    const T av = w == 0 ? tab.alpha0[z * NREF + a]
                        : (T)((double)tab.alpha_w[((size_t)z * NREF + a) * NFREQ + w] * fn.inv[w]);
    acc += gw[atom * NREF + a] * av;
  }
  alpha[t] = acc;
}

template <typename T>
const Tables<T>& tables_of(const d4b200_tables* h);
template <>
const Tables<double>& tables_of<double>(const d4b200_tables* h) { return h->t64; }
template <>
const Tables<float>& tables_of<float>(const d4b200_tables* h) { return h->t32; }

inline unsigned blocks_for(long long total) { return (unsigned)((total + 255) / 256); }

template <typename T>
int weight_references(d4b200_tables_t h, const d4b200_params* par, int nbatch, int nat, const int64_t* numbers,
                      const T* cn, const T* q, T* gw, T* dgwdcn, T* dgwdq, int* status, void* stream) {
  if (!h || !par || nbatch < 0 || nat < 0) return D4B200_EINVAL;
  if (nbatch == 0 || nat == 0) return 0;
  if (!numbers || !gw) return D4B200_EINVAL;
  const bool s = par->model == D4B200_MODEL_D4S;
  const long long total = (long long)nbatch * nat * (s ? nat : 1);
  if (blocks_for(total) == 0 || total > (1ll << 38)) return D4B200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (s)
    k_weight_references<T, true><<<blocks_for(total), 256, 0, st>>>(tables_of<T>(h), h->ga, par->wf, total, nat,
                                                                    numbers, cn, q, gw, dgwdcn, dgwdq, status);
  else
    k_weight_references<T, false><<<blocks_for(total), 256, 0, st>>>(tables_of<T>(h), h->ga, par->wf, total, nat,
                                                                     numbers, cn, q, gw, dgwdcn, dgwdq, status);
  return (int)cudaGetLastError();
}

template <typename T>
int atomic_c6(d4b200_tables_t h, int model, int nbatch, int nat, const int64_t* numbers, const T* gw, T* c6,
              void* stream) {
  if (!h || nbatch < 0 || nat < 0) return D4B200_EINVAL;
  if (nbatch == 0 || nat == 0) return 0;
  if (!numbers || !gw || !c6) return D4B200_EINVAL;
  const long long total = (long long)nbatch * nat * nat;
  cudaStream_t st = (cudaStream_t)stream;
  if (model == D4B200_MODEL_D4S)
    k_atomic_c6<T, true><<<blocks_for(total), 256, 0, st>>>(tables_of<T>(h), total, nat, numbers, gw, c6);
  else
    k_atomic_c6<T, false><<<blocks_for(total), 256, 0, st>>>(tables_of<T>(h), total, nat, numbers, gw, c6);
  return (int)cudaGetLastError();
}

template <typename T>
int weighted_pols(d4b200_tables_t h, int nbatch, int nat, int nfreq, const int64_t* numbers, const T* gw,
                  T* alpha, void* stream) {
  if (!h || nbatch < 0 || nat < 0 || nfreq < 1 || nfreq > NFREQ) return D4B200_EINVAL;
  if (nbatch == 0 || nat == 0) return 0;
  if (!numbers || !gw || !alpha) return D4B200_EINVAL;
  // Casimir-Polder trapezoid weights (src/tad_dftd4/utils.py:52-80)
  static const double cpw[NFREQ] = {2.4999500000000000e-002, 4.9999500000000000e-002, 7.5000000000000010e-002,
                                    0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.15, 0.2, 0.2, 0.2, 0.2, 0.35,
                                    0.5, 0.75, 1.0, 1.75, 2.5, 1.25};
  FreqNorm fn;
  for (int w = 0; w < NFREQ; ++w) fn.inv[w] = 1.0 / sqrt(3.0 / 3.141592653589793238462643383279502884197 * cpw[w]);
  const long long total = (long long)nbatch * nat * nfreq;
  k_weighted_pols<T><<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(tables_of<T>(h), fn, total, nfreq,
                                                                          numbers, gw, alpha);
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" {

int d4b200_weight_references_f64(d4b200_tables_t h, const d4b200_params* par, int nbatch, int nat,
                                 const int64_t* numbers_dev, const double* cn_dev, const double* q_dev,
                                 double* gw_dev, double* dgwdcn_dev, double* dgwdq_dev, int* status_dev,
                                 void* stream) {
  return weight_references<double>(h, par, nbatch, nat, numbers_dev, cn_dev, q_dev, gw_dev, dgwdcn_dev,
                                   dgwdq_dev, status_dev, stream);
}
int d4b200_weight_references_f32(d4b200_tables_t h, const d4b200_params* par, int nbatch, int nat,
                                 const int64_t* numbers_dev, const float* cn_dev, const float* q_dev,
                                 float* gw_dev, float* dgwdcn_dev, float* dgwdq_dev, int* status_dev,
                                 void* stream) {
  return weight_references<float>(h, par, nbatch, nat, numbers_dev, cn_dev, q_dev, gw_dev, dgwdcn_dev,
                                  dgwdq_dev, status_dev, stream);
}
int d4b200_atomic_c6_f64(d4b200_tables_t h, int model, int nbatch, int nat, const int64_t* numbers_dev,
                         const double* gw_dev, double* c6_dev, void* stream) {
  return atomic_c6<double>(h, model, nbatch, nat, numbers_dev, gw_dev, c6_dev, stream);
}
int d4b200_atomic_c6_f32(d4b200_tables_t h, int model, int nbatch, int nat, const int64_t* numbers_dev,
                         const float* gw_dev, float* c6_dev, void* stream) {
  return atomic_c6<float>(h, model, nbatch, nat, numbers_dev, gw_dev, c6_dev, stream);
}
int d4b200_weighted_pols_f64(d4b200_tables_t h, int nbatch, int nat, int nfreq, const int64_t* numbers_dev,
                             const double* gw_dev, double* alpha_dev, void* stream) {
  return weighted_pols<double>(h, nbatch, nat, nfreq, numbers_dev, gw_dev, alpha_dev, stream);
}
int d4b200_weighted_pols_f32(d4b200_tables_t h, int nbatch, int nat, int nfreq, const int64_t* numbers_dev,
                             const float* gw_dev, float* alpha_dev, void* stream) {
  return weighted_pols<float>(h, nbatch, nat, nfreq, numbers_dev, gw_dev, alpha_dev, stream);
}

}  // extern "C"

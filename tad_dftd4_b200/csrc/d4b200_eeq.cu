// EEQ-2019 atomic partial charges for padded batches of small structures, and the
// vector-Jacobian product of the charges with respect to the positions.
//
// Replaces tad_multicharge.get_eeq_charges (third-party, tad-multicharge==0.5.0;
// call sites /root/reference/src/tad_dftd4/dispersion/base.py:401-407 and disp.py:190),
// the step immediately before the D4 hot path.
//
// One CTA per structure, everything in shared memory:
//   compact real atoms -> erf coordination number (near pairs only) -> bordered matrix
//   [[A, 1], [1^T, 0]] with A_ij = erf(gamma_ij r_ij)/r_ij built once per unordered pair ->
//   right-looking elimination of the upper triangle of [M | rhs] without pivoting (A is the
//   Gram matrix of Gaussian charges plus the hardness: positive definite for physical
//   geometries; the border pivot is -1^T A^-1 1) -> back substitution by one warp.
// VJP kernel: same matrix, rhs = (dL/dq, 0) -> mu; then
//   dL/dx = mu^T (d rhs/dx - dM/dx (q, lambda))
// as one weight per unordered pair (stored in the dead matrix) and a row sweep.
// The arithmetic is float64 for both I/O types (the system is ill-suited to float32 and
// costs a few per cent of the D4 kernels).  Model in numpy: tests/eeq_model.py.

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/d4b200.h"

namespace {

constexpr int EEQ_NELEM = 87;  // Z = 0 (padding) .. 86
constexpr int EEQ_NT_SMALL = 128;  // structures padded to <= 64 atoms (several CTAs per SM)
constexpr int EEQ_NT_LARGE = 512;  // beyond: one or two CTAs per SM, more warps per elimination step
constexpr int EEQ_MAX_NAT = 160;  // (14 (nat+1) + (nat+1) ld) doubles <= 227 KB
constexpr int EEQ_NARR = 14;

constexpr double KCN = 7.5;
constexpr double CN_MAX = 8.0;
constexpr double LOG1P_EXP_CN_MAX = 8.0003354063728956;  // log(1 + exp(8))
constexpr double SQRT_2_OVER_PI = 0.79788456080286535588;
constexpr double TWO_OVER_SQRT_PI = 1.12837916709551257390;
constexpr double ONE_OVER_SQRT_PI = 0.56418958354775628695;
constexpr double DBL_EPS = 2.220446049250313e-16;

struct EeqTables {
  const double *chi, *eta, *kcn, *rad, *rcov;
};

__host__ __device__ inline int eeq_ld(int nat) { return (nat + 2) | 1; }
// doubles per structure of the saved factor: matrix rows, reciprocal pivots, raw coordination numbers
__host__ inline size_t eeq_factor_doubles(int nat) { return (size_t)(nat + 1) * eeq_ld(nat) + 2 * (size_t)(nat + 1); }
__host__ inline size_t eeq_smem(int nat) {
  return sizeof(double) * ((size_t)EEQ_NARR * (nat + 1) + (size_t)(nat + 1) * eeq_ld(nat));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// pair index p -> (i, j), i > j >= 0, p = i (i - 1) / 2 + j
__device__ __forceinline__ void pair_of(int p, int& i, int& j) {
  i = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
  while (i * (i - 1) / 2 > p) --i;
  while ((i + 1) * i / 2 <= p) ++i;
  j = p - i * (i - 1) / 2;
}

template <typename T, bool VJP, int EEQ_NT>
__global__ void __launch_bounds__(EEQ_NT)
eeq_kernel(EeqTables tb, int nat, const int64_t* __restrict__ numbers, const T* __restrict__ pos,
           const T* __restrict__ charge, double cutoff2, const T* __restrict__ q_in,
           const T* __restrict__ gq, T* __restrict__ out, int* __restrict__ status,
           double* __restrict__ fac, size_t fac_stride) {
  extern __shared__ double sm[];
  const int b = blockIdx.x;
  constexpr int NW = EEQ_NT / 32, EEQ_NW = NW;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int na1 = nat + 1;
  double* xs = sm;
  double* ys = xs + na1;
  double* zs = ys + na1;
  double* rad = zs + na1;
  double* rcv = rad + na1;
  double* kap = rcv + na1;
  double* cnr = kap + na1;   // raw coordination number; later force x
  double* cnc = cnr + na1;   // cut coordination number; later force y
  double* sol = cnc + na1;   // solution of the bordered system
  double* dinv = sol + na1;  // reciprocal pivots; later force z
  double* qv = dinv + na1;   // VJP: charges
  double* big = qv + na1;    // VJP: dL/dcn_raw
  int* zat = reinterpret_cast<int*>(big + na1);  // [nat] atomic number by padded slot
  int* idx = zat + na1;                          // [nat] compact -> padded slot
  int* inv = idx + na1;                          // [nat] padded slot -> compact (-1 = padding)
  int* cnt = inv + na1;                          // [1]
  double* M = sm + (size_t)EEQ_NARR * na1;

  const int64_t* zrow = numbers + (size_t)b * nat;
  for (int a = tid; a < nat; a += EEQ_NT) {
    long long z = zrow[a];
    if (z < 0 || z >= EEQ_NELEM) {
      if (status) atomicOr(status, D4B200_STATUS_BAD_NUMBER);
      z = 0;
    }
    zat[a] = (int)z;
  }
  __syncthreads();
  if (warp == 0) {
    int n = 0;
    for (int base = 0; base < nat; base += 32) {
      const int a = base + lane;
      const bool real = a < nat && zat[a] != 0;
      const unsigned m = __ballot_sync(0xffffffffu, real);
      const int c = n + __popc(m & ((1u << lane) - 1u));
      if (a < nat) inv[a] = real ? c : -1;
      if (real) idx[c] = a;
      n += __popc(m);
    }
    if (lane == 0) *cnt = n;
  }
  __syncthreads();
  const int n = *cnt;
  const int m = n + 1;       // bordered system size; the rhs is column m
  const int ld = (m + 1) | 1;  // odd stride: conflict-free column walks
  T* orow = out + (size_t)b * nat * (VJP ? 3 : 1);
  if (n == 0) {
    for (int a = tid; a < nat * (VJP ? 3 : 1); a += EEQ_NT) orow[a] = T(0);
    return;
  }
  for (int c = tid; c < n; c += EEQ_NT) {
    const int a = idx[c];
    const int z = zat[a];
    const T* p = pos + ((size_t)b * nat + a) * 3;
    xs[c] = (double)p[0];
    ys[c] = (double)p[1];
    zs[c] = (double)p[2];
    rad[c] = tb.rad[z];
    rcv[c] = tb.rcov[z];
    kap[c] = tb.kcn[z];
    if (VJP) qv[c] = (double)q_in[(size_t)b * nat + a];
  }
  __syncthreads();

  // The factor of the bordered matrix does not depend on the right-hand side: the charges kernel can
  // leave it (upper triangle after the elimination, reciprocal pivots, raw coordination numbers) in
  // global memory, and the VJP kernel of the same geometry then only substitutes its own right-hand side.
  double* const fb = fac ? fac + (size_t)b * fac_stride : nullptr;
  const size_t fac_dinv = (size_t)(nat + 1) * eeq_ld(nat), fac_cn = fac_dinv + (nat + 1);
  const bool reuse = VJP && fb != nullptr;
  const int npair = n * (n - 1) / 2;
  if (!reuse) {
  // ---- coordination number: ordered rows; the few near pairs of a row (r < 1.8 r0, beyond
  // which erfc/2 < 1e-17) are first compacted into a per-warp list (in the not yet used matrix
  // area) so that the expensive erfc runs once per 32 near pairs, not once per 32 pairs.
  {
    int* near = reinterpret_cast<int*>(M) + warp * n;  // <= n - 1 entries; 4 NW n bytes fit the matrix
    for (int i = warp; i < n; i += EEQ_NW) {
      const double xi = xs[i], yi = ys[i], zi = zs[i], ri = rcv[i];
      int cntn = 0;
      for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        bool isnear = false;
        if (j < n && j != i) {
          const double dx = xi - xs[j], dy = yi - ys[j], dz = zi - zs[j];
          const double r2 = dx * dx + dy * dy + dz * dz;
          const double r0 = ri + rcv[j];
          isnear = r2 <= cutoff2 && r2 < 3.24 * r0 * r0;
        }
        const unsigned mk = __ballot_sync(0xffffffffu, isnear);
        if (isnear) near[cntn + __popc(mk & ((1u << lane) - 1u))] = j;
        cntn += __popc(mk);
      }
      __syncwarp();
      double acc = 0.0;
      for (int e = lane; e < cntn; e += 32) {
        const int j = near[e];
        const double dx = xi - xs[j], dy = yi - ys[j], dz = zi - zs[j];
        const double r2 = dx * dx + dy * dy + dz * dz;
        acc += 0.5 * erfc(KCN * (sqrt(r2) / (ri + rcv[j]) - 1.0));
      }
      acc = warp_sum(acc);
      if (lane == 0) cnr[i] = acc;
      __syncwarp();
    }
  }
  __syncthreads();  // the lists lived in the matrix area
  // smooth cut at CN_MAX: log(1 + e^8) - log(1 + e^(8 - cn)), one atom per thread
  for (int c = tid; c < n; c += EEQ_NT) cnc[c] = LOG1P_EXP_CN_MAX - log1p(exp(CN_MAX - cnr[c]));

  // ---- Coulomb block, once per unordered pair --------------------------------------------
  for (int p = tid; p < npair; p += EEQ_NT) {
    int i, j;
    pair_of(p, i, j);
    const double dx = xs[i] - xs[j], dy = ys[i] - ys[j], dz = zs[i] - zs[j];
    const double r2 = dx * dx + dy * dy + dz * dz;
    const double rinv = rsqrt(r2);
    const double g = rsqrt(rad[i] * rad[i] + rad[j] * rad[j]);
    M[j * ld + i] = erf(g * (r2 * rinv)) * rinv;  // upper triangle (j < i)
  }
  __syncthreads();
  for (int c = tid; c < n; c += EEQ_NT) {
    const int a = idx[c];
    const int z = zat[a];
    const double dg = tb.eta[z] + SQRT_2_OVER_PI / rad[c];
    M[c * ld + c] = dg;
    if (c == 0) dinv[0] = 1.0 / dg;
    M[c * ld + n] = 1.0;
    M[c * ld + m] = VJP ? (double)gq[(size_t)b * nat + a]
                        : -tb.chi[z] + kap[c] * sqrt(fmax(cnc[c], DBL_EPS));
  }
  if (tid == 0) {
    M[n * ld + n] = 0.0;
    M[n * ld + m] = VJP ? 0.0 : (double)charge[b];
  }
  __syncthreads();

  // ---- elimination of [M | rhs], no pivoting ----------------------------------------------
  // The trailing matrix stays symmetric, so only its upper triangle (j >= i) and the rhs are
  // updated, and the multiplier of row i comes from the pivot ROW: f_i = a_ki / a_kk.  One
  // block barrier per column.  A warp updates four of its rows at a time (loads grouped before
  // the stores, so the shared-memory latencies overlap); the reciprocal of the next pivot is
  // produced by the lane that has just updated it, off the other warps' path.
  for (int k = 0; k < m; ++k) {
    const double pinv = dinv[k];
    const double* rk = M + k * ld;
    // blocks of four ADJACENT rows (their upper-triangle column ranges nearly coincide)
    for (int i0 = k + 1 + 4 * warp; i0 < m; i0 += 4 * NW) {
      double* r0 = M + i0 * ld;
      const bool v1 = i0 + 1 < m, v2 = i0 + 2 < m, v3 = i0 + 3 < m;
      double* r1 = v1 ? r0 + ld : r0;
      double* r2 = v2 ? r0 + 2 * ld : r0;
      double* r3 = v3 ? r0 + 3 * ld : r0;
      const double f0 = rk[i0] * pinv, f1 = v1 ? rk[i0 + 1] * pinv : 0.0,
                   f2 = v2 ? rk[i0 + 2] * pinv : 0.0, f3 = v3 ? rk[i0 + 3] * pinv : 0.0;
      for (int j = i0 + lane; j <= m; j += 32) {  // up to three columns left of a diagonal: harmless
        const double p = rk[j];
        double a0 = r0[j], a1 = r1[j], a2 = r2[j], a3 = r3[j];
        a0 = fma(-f0, p, a0);
        a1 = fma(-f1, p, a1);
        a2 = fma(-f2, p, a2);
        a3 = fma(-f3, p, a3);
        r0[j] = a0;
        if (v1) r1[j] = a1;
        if (v2) r2[j] = a2;
        if (v3) r3[j] = a3;
      }
      if (i0 == k + 1 && lane == 0) dinv[k + 1] = 1.0 / r0[k + 1];
    }
    __syncthreads();
  }
  if (!VJP && fb) {
    for (int i = warp; i < m; i += EEQ_NW)
      for (int j = i + lane; j < m; j += 32) fb[(size_t)i * ld + j] = M[i * ld + j];
    for (int c = tid; c < m; c += EEQ_NT) fb[fac_dinv + c] = dinv[c];
    for (int c = tid; c < n; c += EEQ_NT) fb[fac_cn + c] = cnr[c];
  }
  } else {
    for (int i = warp; i < m; i += EEQ_NW)
      for (int j = i + lane; j < m; j += 32) M[i * ld + j] = fb[(size_t)i * ld + j];
    for (int c = tid; c < m; c += EEQ_NT) dinv[c] = fb[fac_dinv + c];
    for (int c = tid; c < n; c += EEQ_NT) {
      const double cn = fb[fac_cn + c];
      cnr[c] = cn;
      cnc[c] = LOG1P_EXP_CN_MAX - log1p(exp(CN_MAX - cn));
      M[c * ld + m] = (double)gq[(size_t)b * nat + idx[c]];
    }
    if (tid == 0) M[n * ld + m] = 0.0;
    __syncthreads();
  }
  // ---- back substitution: warp 0, right-hand side in registers, no block barriers -----------
  if (warp == 0) {
    constexpr int MAXC = (EEQ_MAX_NAT + 1 + 31) / 32;
    double bb[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int i = lane + 32 * c;
      bb[c] = i < m ? M[i * ld + m] : 0.0;
    }
    if (reuse) {  // the elimination steps applied to this right-hand side: b_i -= (u_ki / u_kk) b_k, i > k
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (32 * c < m) {
          for (int kk = 0; kk <= min(31, m - 1 - 32 * c); ++kk) {
            const int k = 32 * c + kk;
            const double xk = __shfl_sync(0xffffffffu, bb[c], kk) * dinv[k];
#pragma unroll
            for (int cc = c; cc < MAXC; ++cc) {
              const int i = lane + 32 * cc;
              if (i > k && i < m) bb[cc] = fma(-M[k * ld + i], xk, bb[cc]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int c = MAXC - 1; c >= 0; --c) {
      if (32 * c < m) {
        for (int kk = min(31, m - 1 - 32 * c); kk >= 0; --kk) {
          const int k = 32 * c + kk;
          const double xk = __shfl_sync(0xffffffffu, bb[c], kk) * dinv[k];
          if (lane == kk) sol[k] = xk;
#pragma unroll
          for (int cc = 0; cc <= c; ++cc) {
            const int i = lane + 32 * cc;
            if (i < k) bb[cc] = fma(-M[i * ld + k], xk, bb[cc]);
          }
        }
      }
    }
  }
  __syncthreads();

  if constexpr (!VJP) {
    for (int a = tid; a < nat; a += EEQ_NT) {
      const int c = inv[a];
      orow[a] = c >= 0 ? (T)sol[c] : T(0);
    }
  } else {
  // ---- vector-Jacobian product ---------------------------------------------------------------
  for (int c = tid; c < n; c += EEQ_NT) {
    const double cc = cnc[c];
    const double dsq = cc > DBL_EPS ? 0.5 * sol[c] * kap[c] / sqrt(cc) : 0.0;
    big[c] = dsq / (1.0 + exp(cnr[c] - CN_MAX));
  }
  __syncthreads();
  for (int p = tid; p < npair; p += EEQ_NT) {
    int i, j;
    pair_of(p, i, j);
    const double dx = xs[i] - xs[j], dy = ys[i] - ys[j], dz = zs[i] - zs[j];
    const double r2 = dx * dx + dy * dy + dz * dz;
    const double rinv = rsqrt(r2);
    const double r = r2 * rinv;
    const double g = rsqrt(rad[i] * rad[i] + rad[j] * rad[j]);
    const double da = (TWO_OVER_SQRT_PI * g * exp(-g * g * r2) - erf(g * r) * rinv) * rinv;
    double w = -(sol[i] * qv[j] + sol[j] * qv[i]) * da;
    const double r0 = rcv[i] + rcv[j];
    const double za = KCN * (r / r0 - 1.0);
    if (r2 <= cutoff2 && za * za < 40.0) {
      w -= (big[i] + big[j]) * (KCN * ONE_OVER_SQRT_PI / r0) * exp(-za * za);
    }
    w *= rinv;
    M[i * ld + j] = w;
    M[j * ld + i] = w;
  }
  __syncthreads();
  double* fx = cnr;
  double* fy = cnc;
  double* fz = dinv;
  for (int i = warp; i < n; i += EEQ_NW) {
    const double xi = xs[i], yi = ys[i], zi = zs[i];
    const double* wi = M + i * ld;
    double ax = 0.0, ay = 0.0, az = 0.0;
    for (int j = lane; j < n; j += 32) {
      if (j != i) {
        const double w = wi[j];
        ax = fma(w, xi - xs[j], ax);
        ay = fma(w, yi - ys[j], ay);
        az = fma(w, zi - zs[j], az);
      }
    }
    ax = warp_sum(ax);
    ay = warp_sum(ay);
    az = warp_sum(az);
    if (lane == 0) {
      fx[i] = ax;
      fy[i] = ay;
      fz[i] = az;
    }
  }
  __syncthreads();
  for (int a = tid; a < nat; a += EEQ_NT) {
    const int c = inv[a];
    orow[a * 3 + 0] = c >= 0 ? (T)fx[c] : T(0);
    orow[a * 3 + 1] = c >= 0 ? (T)fy[c] : T(0);
    orow[a * 3 + 2] = c >= 0 ? (T)fz[c] : T(0);
  }
  }  // VJP
}

}  // namespace

struct d4b200_eeq {
  int device;
  double* blob;
  EeqTables t;
};

static thread_local long long g_eeq_launches = 0;

extern "C" {

int d4b200_eeq_limit(void) { return EEQ_MAX_NAT; }
long long d4b200_eeq_launch_count(void) { return g_eeq_launches; }

int d4b200_eeq_create(int device, const double* param_host, size_t n, d4b200_eeq_t* out) {
  if (!param_host || !out) return D4B200_EINVAL;
  if (n != (size_t)5 * EEQ_NELEM) return D4B200_ETABLE;
  int prev = 0;
  cudaError_t e = cudaGetDevice(&prev);
  if (e != cudaSuccess) return (int)e;
  if ((e = cudaSetDevice(device)) != cudaSuccess) return (int)e;
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return (int)e;
  if (prop.major != 10) {
    cudaSetDevice(prev);
    return D4B200_EARCH;
  }
  d4b200_eeq* h = new d4b200_eeq();
  h->device = device;
  if ((e = cudaMalloc(&h->blob, n * sizeof(double))) != cudaSuccess) {
    delete h;
    cudaSetDevice(prev);
    return (int)e;
  }
  e = cudaMemcpy(h->blob, param_host, n * sizeof(double), cudaMemcpyHostToDevice);
  h->t.chi = h->blob;
  h->t.eta = h->blob + EEQ_NELEM;
  h->t.kcn = h->blob + 2 * EEQ_NELEM;
  h->t.rad = h->blob + 3 * EEQ_NELEM;
  h->t.rcov = h->blob + 4 * EEQ_NELEM;
  const int smem = (int)eeq_smem(EEQ_MAX_NAT);
  auto attr = [&](auto kern) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  };
  attr(eeq_kernel<double, false, EEQ_NT_SMALL>);
  attr(eeq_kernel<double, true, EEQ_NT_SMALL>);
  attr(eeq_kernel<float, false, EEQ_NT_SMALL>);
  attr(eeq_kernel<float, true, EEQ_NT_SMALL>);
  attr(eeq_kernel<double, false, EEQ_NT_LARGE>);
  attr(eeq_kernel<double, true, EEQ_NT_LARGE>);
  attr(eeq_kernel<float, false, EEQ_NT_LARGE>);
  attr(eeq_kernel<float, true, EEQ_NT_LARGE>);
  cudaSetDevice(prev);
  if (e != cudaSuccess) {
    cudaFree(h->blob);
    delete h;
    return (int)e;
  }
  *out = h;
  return 0;
}

int d4b200_eeq_destroy(d4b200_eeq_t h) {
  if (!h) return D4B200_EINVAL;
  cudaFree(h->blob);
  delete h;
  return 0;
}

}  // extern "C"

template <typename T, bool VJP>
static int eeq_launch(d4b200_eeq_t h, int nbatch, int nat, const int64_t* numbers, const T* pos,
                      const T* charge, double cutoff, const T* q, const T* gq, T* out, int* status,
                      void* stream, double* factor = nullptr) {
  if (!h || nbatch < 0 || nat < 0) return D4B200_EINVAL;
  if (nbatch == 0 || nat == 0) return 0;
  if (!numbers || !pos || !out || (VJP ? (!q || !gq) : !charge)) return D4B200_EINVAL;
  if (nat > EEQ_MAX_NAT) return D4B200_EINVAL;
  if (!(cutoff > 0.0)) return D4B200_EPARAM;
  if (nat <= 64)
    eeq_kernel<T, VJP, EEQ_NT_SMALL><<<nbatch, EEQ_NT_SMALL, eeq_smem(nat), (cudaStream_t)stream>>>(
        h->t, nat, numbers, pos, charge, cutoff * cutoff, q, gq, out, status, factor, eeq_factor_doubles(nat));
  else
    eeq_kernel<T, VJP, EEQ_NT_LARGE><<<nbatch, EEQ_NT_LARGE, eeq_smem(nat), (cudaStream_t)stream>>>(
        h->t, nat, numbers, pos, charge, cutoff * cutoff, q, gq, out, status, factor, eeq_factor_doubles(nat));
  ++g_eeq_launches;
  return (int)cudaGetLastError();
}

extern "C" {

int d4b200_eeq_charges_f64(d4b200_eeq_t h, int nbatch, int nat, const int64_t* numbers_dev,
                           const double* positions_dev, const double* charge_dev, double cn_cutoff,
                           double* q_dev, int* status_dev, void* stream) {
  return eeq_launch<double, false>(h, nbatch, nat, numbers_dev, positions_dev, charge_dev, cn_cutoff,
                                   nullptr, nullptr, q_dev, status_dev, stream);
}
int d4b200_eeq_charges_f32(d4b200_eeq_t h, int nbatch, int nat, const int64_t* numbers_dev,
                           const float* positions_dev, const float* charge_dev, double cn_cutoff,
                           float* q_dev, int* status_dev, void* stream) {
  return eeq_launch<float, false>(h, nbatch, nat, numbers_dev, positions_dev, charge_dev, cn_cutoff,
                                  nullptr, nullptr, q_dev, status_dev, stream);
}
int d4b200_eeq_vjp_f64(d4b200_eeq_t h, int nbatch, int nat, const int64_t* numbers_dev,
                       const double* positions_dev, double cn_cutoff, const double* q_dev,
                       const double* grad_q_dev, double* grad_positions_dev, int* status_dev,
                       void* stream) {
  return eeq_launch<double, true>(h, nbatch, nat, numbers_dev, positions_dev, nullptr, cn_cutoff, q_dev,
                                  grad_q_dev, grad_positions_dev, status_dev, stream);
}
int d4b200_eeq_vjp_f32(d4b200_eeq_t h, int nbatch, int nat, const int64_t* numbers_dev,
                       const float* positions_dev, double cn_cutoff, const float* q_dev,
                       const float* grad_q_dev, float* grad_positions_dev, int* status_dev,
                       void* stream) {
  return eeq_launch<float, true>(h, nbatch, nat, numbers_dev, positions_dev, nullptr, cn_cutoff, q_dev,
                                 grad_q_dev, grad_positions_dev, status_dev, stream);
}


// Charges + saved factor / VJP from the saved factor (same geometry): the elimination is not repeated in the
// backward pass.  factor_dev: nbatch * d4b200_eeq_factor_doubles(nat) doubles (may be NULL: plain call).
size_t d4b200_eeq_factor_doubles(int nat) { return nat >= 0 && nat <= EEQ_MAX_NAT ? eeq_factor_doubles(nat) : 0; }
int d4b200_eeq_charges_factor_f64(d4b200_eeq_t h, int nbatch, int nat, const int64_t* numbers_dev,
                                  const double* positions_dev, const double* charge_dev, double cn_cutoff,
                                  double* q_dev, double* factor_dev, int* status_dev, void* stream) {
  return eeq_launch<double, false>(h, nbatch, nat, numbers_dev, positions_dev, charge_dev, cn_cutoff,
                                   nullptr, nullptr, q_dev, status_dev, stream, factor_dev);
}
int d4b200_eeq_charges_factor_f32(d4b200_eeq_t h, int nbatch, int nat, const int64_t* numbers_dev,
                                  const float* positions_dev, const float* charge_dev, double cn_cutoff,
                                  float* q_dev, double* factor_dev, int* status_dev, void* stream) {
  return eeq_launch<float, false>(h, nbatch, nat, numbers_dev, positions_dev, charge_dev, cn_cutoff,
                                  nullptr, nullptr, q_dev, status_dev, stream, factor_dev);
}
int d4b200_eeq_vjp_factor_f64(d4b200_eeq_t h, int nbatch, int nat, const int64_t* numbers_dev,
                              const double* positions_dev, double cn_cutoff, const double* q_dev,
                              const double* grad_q_dev, const double* factor_dev, double* grad_positions_dev,
                              int* status_dev, void* stream) {
  return eeq_launch<double, true>(h, nbatch, nat, numbers_dev, positions_dev, nullptr, cn_cutoff, q_dev,
                                  grad_q_dev, grad_positions_dev, status_dev, stream,
                                  const_cast<double*>(factor_dev));
}
int d4b200_eeq_vjp_factor_f32(d4b200_eeq_t h, int nbatch, int nat, const int64_t* numbers_dev,
                              const float* positions_dev, double cn_cutoff, const float* q_dev,
                              const float* grad_q_dev, const double* factor_dev, float* grad_positions_dev,
                              int* status_dev, void* stream) {
  return eeq_launch<float, true>(h, nbatch, nat, numbers_dev, positions_dev, nullptr, cn_cutoff, q_dev,
                                 grad_q_dev, grad_positions_dev, status_dev, stream,
                                 const_cast<double*>(factor_dev));
}

}  // extern "C"

// Vector-Jacobian product of the atom-resolved D4 energy with respect to the seven damping
// parameters (s6, s8, s9, s10, a1, a2, alp): what the reference obtains by differentiating
// its dense tape (test/test_grad/test_param.py:40-100; Param values are 0-d tensors,
// src/tad_dftd4/damping/parameters/base.py:48-85).
//
//   L = sum_i g_i E_i,  E = E2 + E3
//   E2_i = -1/2 sum_j C6q_ij [ s6 t6 + s8 Q t8 + s10 (49/40) Q^2 t10 ],  t_n = 1/(r^n + R0^n)
//        (src/tad_dftd4/dispersion/twobody.py:134-201, damping/functions.py:262-305)
//   E3_i = 1/6 sum_jk M_ijk (0.375 s / r^5 + 1 / r^3) s9 sqrt|C60 C60 C60| / (1 + 6 (R0/r)^(alp/3))
//        (dispersion/threebody.py:54-163, 244-256, 311-321)
//   R0_ij = a1 sqrt(3 r4r2_i r4r2_j) + a2
//
// L is linear in s6, s8, s9, s10.  a1 and a2 enter the two-body term through
// d t_n / d R0 = -n R0^(n-1) t_n^2 and, together with alp, the ATM damping through the per-pair
// factors u_p = (R0_p / r_p)^(alp/3): with t = u_ij u_ik u_jk the triple energy obeys
// d e / d ln u_p = -6 t e / (1 + 6 t) for each of its three pairs.  Algebra checked on the
// CPU in tests/kernel_model.py::param_gradient against autograd of the oracle.
//
// Inputs are the pair C6 matrices of both flavours (charge-scaled for the two-body term,
// q = 0 for ATM), produced on device by d4b200_weight_references_* / d4b200_atomic_c6_*, so
// the same entry point serves the D4 and the D4S model.  Three launches:
//   k_param_pairs    CTA per (structure, atom i): two-body sums of row i, ATM pair planes
//   k_param_triples  CTA per (structure, top atom i): all triples i > j > k, once each
//   k_param_reduce   CTA per structure: fixed-order sum over the rows -> out[b][7]
// Sums are formed in a fixed order (bitwise reproducible); accumulation is float64 for both
// input types.  This path serves parameter fitting on small molecules; it is not a
// throughput path.
#include <cuda_runtime.h>
#include <math.h>

#include "d4b200_handle.cuh"

namespace {

using namespace d4b200;

constexpr int NPLANE = 6;  // signed r^2, sqrt|C60|/r^5, u, ss/R0, 1/R0, ln(R0/r)
constexpr int N2 = 5;      // two-body partial sums per row: s6, s8, s10, a1, a2
constexpr int N3 = 4;      // ATM partial sums per row: s9, a1, a2, alp

struct ParV {
  double s6, s8, s9, s10k, a1, a2, alp3;
  double disp2_sq, disp3_sq;
  int has_s10;
};

template <int K, int NT>
__device__ __forceinline__ void block_sum(double (&v)[K], double* out) {
  __shared__ double red[K][NT / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double s = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[k][warp] = s;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double s = 0.0;
    for (int w = 0; w < NT / 32; ++w) s += red[threadIdx.x][w];
    out[threadIdx.x] = s;
  }
}

template <typename T>
__global__ void __launch_bounds__(128) k_param_pairs(Tables<T> tab, ParV P, int nat,
                                                     const int64_t* __restrict__ numbers,
                                                     const T* __restrict__ pos, const T* __restrict__ c6q,
                                                     const T* __restrict__ c60, const T* __restrict__ gin,
                                                     double* __restrict__ planes, double* __restrict__ part2) {
  const int i = blockIdx.x, b = blockIdx.y;
  const size_t row = (size_t)b * nat + i;
  const size_t plane = (size_t)gridDim.y * nat * nat;
  long long zi = numbers[row];
  if (zi < 0 || zi >= NELEM) zi = 0;
  double acc[N2] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (zi != 0) {
    const double xi = (double)pos[3 * row], yi = (double)pos[3 * row + 1], zc = (double)pos[3 * row + 2];
    const double sqi = (double)tab.sqrt_r4r2[zi];
    const double gi = gin ? (double)gin[row] : 1.0;
    for (int j = threadIdx.x; j < nat; j += 128) {
      const size_t col = (size_t)b * nat + j;
      long long zj = numbers[col];
      if (zj <= 0 || zj >= NELEM || j == i) continue;
      const double dx = xi - (double)pos[3 * col], dy = yi - (double)pos[3 * col + 1],
                   dz = zc - (double)pos[3 * col + 2];
      const double r2 = dx * dx + dy * dy + dz * dz;
      const double r = sqrt(r2), rinv = 1.0 / r;
      const double ss = sqi * (double)tab.sqrt_r4r2[zj];  // sqrt(3 r4r2_i r4r2_j)
      const double R0 = P.a1 * ss + P.a2;
      const size_t o = row * nat + j;
      if (r2 <= P.disp2_sq) {
        const double gj = gin ? (double)gin[col] : 1.0;
        const double w = -0.25 * (gi + gj) * (double)c6q[o];  // every unordered pair is met twice
        const double qq = ss * ss;
        const double r4 = r2 * r2, r6 = r4 * r2, r8 = r4 * r4;
        const double R2 = R0 * R0, R4 = R2 * R2, R5 = R4 * R0, R6 = R4 * R2, R8 = R4 * R4;
        const double t6 = 1.0 / (r6 + R6), t8 = 1.0 / (r8 + R8);
        double dF = 6.0 * P.s6 * R5 * t6 * t6 + 8.0 * P.s8 * qq * (R6 * R0) * t8 * t8;
        acc[0] += w * t6;
        acc[1] += w * qq * t8;
        if (P.has_s10) {
          const double t10 = 1.0 / (r8 * r2 + R8 * R2);
          acc[2] += w * (49.0 / 40.0) * qq * qq * t10;
          dF += 10.0 * P.s10k * qq * qq * (R8 * R0) * t10 * t10;
        }
        acc[3] -= w * dF * ss;
        acc[4] -= w * dF;
      }
      const double x = R0 * rinv;
      const double lg = log(x);
      const double ri2 = rinv * rinv;
      planes[o] = r2 <= P.disp3_sq ? r2 : -r2;
      planes[plane + o] = sqrt(fabs((double)c60[o])) * (ri2 * ri2 * rinv);
      planes[2 * plane + o] = exp(P.alp3 * lg);
      planes[3 * plane + o] = ss / R0;
      planes[4 * plane + o] = 1.0 / R0;
      planes[5 * plane + o] = lg;
    }
  }
  block_sum<N2, 128>(acc, part2 + row * N2);
}

template <typename T>
__global__ void __launch_bounds__(256) k_param_triples(ParV P, int nat, const int64_t* __restrict__ numbers,
                                                       const T* __restrict__ gin,
                                                       const double* __restrict__ planes,
                                                       double* __restrict__ part3) {
  const int i = blockIdx.x, b = blockIdx.y;
  const size_t row = (size_t)b * nat + i;
  const size_t plane = (size_t)gridDim.y * nat * nat;
  const int64_t* zrow = numbers + (size_t)b * nat;
  double acc[N3] = {0.0, 0.0, 0.0, 0.0};
  if (zrow[i] != 0) {
    const double gi = gin ? (double)gin[row] : 1.0;
    const double* const ri = planes + row * nat;  // entries (i, .)
    const long long np = (long long)i * (i - 1) / 2;
    for (long long p = threadIdx.x; p < np; p += 256) {
      // p = j (j - 1) / 2 + k, k < j < i
      int j = (int)((1.0 + sqrt(1.0 + 8.0 * (double)p)) * 0.5);
      while ((long long)j * (j - 1) / 2 > p) --j;
      while ((long long)(j + 1) * j / 2 <= p) ++j;
      const int k = (int)(p - (long long)j * (j - 1) / 2);
      if (zrow[j] == 0 || zrow[k] == 0) continue;
      const double* const rj = planes + ((size_t)b * nat + j) * nat;  // entries (j, .)
      const double as = ri[j], cs = ri[k], bs = rj[k];
      const double cij = as > 0.0 ? 1.0 : 0.0, cik = cs > 0.0 ? 1.0 : 0.0, cjk = bs > 0.0 ? 1.0 : 0.0;
      const double gj = gin ? (double)gin[(size_t)b * nat + j] : 1.0;
      const double gk = gin ? (double)gin[(size_t)b * nat + k] : 1.0;
      // threebody.py:153-157 tests r_ij and r_jk only -> per-atom multiplicities
      const double W = gi * cjk * (cij + cik) + gj * cik * (cij + cjk) + gk * cij * (cik + cjk);
      if (W == 0.0) continue;
      const double a = fabs(as), c = fabs(cs), bb = fabs(bs);
      const double t1 = a - c;
      const double s = (bb * bb - t1 * t1) * (a + c - bb);
      const double t = ri[2 * plane + j] * ri[2 * plane + k] * rj[2 * plane + k];
      const double f = 1.0 / (1.0 + 6.0 * t);
      const double pp = ri[plane + j] * ri[plane + k] * rj[plane + k];
      const double e = W * (0.375 * s + a * bb * c) * pp * f * (1.0 / 6.0);
      const double h = -6.0 * t * f * e * P.s9;
      acc[0] += e;
      acc[1] += h * P.alp3 * (ri[3 * plane + j] + ri[3 * plane + k] + rj[3 * plane + k]);
      acc[2] += h * P.alp3 * (ri[4 * plane + j] + ri[4 * plane + k] + rj[4 * plane + k]);
      acc[3] += h * (1.0 / 3.0) * (ri[5 * plane + j] + ri[5 * plane + k] + rj[5 * plane + k]);
    }
  }
  block_sum<N3, 256>(acc, part3 + row * N3);
}

// out[b] = (s6, s8, s9, s10, a1, a2, alp), summed over the rows in index order
__global__ void k_param_reduce(int nat, const double* __restrict__ part2, const double* __restrict__ part3,
                               double* __restrict__ out) {
  const int b = blockIdx.x, k = threadIdx.x;
  if (k >= 7) return;
  double s = 0.0;
  for (int i = 0; i < nat; ++i) {
    const double* p2 = part2 + ((size_t)b * nat + i) * N2;
    const double* p3 = part3 + ((size_t)b * nat + i) * N3;
    switch (k) {
      case 0: s += p2[0]; break;
      case 1: s += p2[1]; break;
      case 2: s += p3[0]; break;
      case 3: s += p2[2]; break;
      case 4: s += p2[3] + p3[1]; break;
      case 5: s += p2[4] + p3[2]; break;
      default: s += p3[3]; break;
    }
  }
  out[(size_t)b * 7 + k] = s;
}

size_t ws_bytes(int nbatch, int nat) {
  const size_t pairs = (size_t)nbatch * nat * nat, rows = (size_t)nbatch * nat;
  return (NPLANE * pairs + (N2 + N3) * rows) * sizeof(double);
}

template <typename T>
const Tables<T>& tables_of(const d4b200_tables* h);
template <>
const Tables<double>& tables_of<double>(const d4b200_tables* h) { return h->t64; }
template <>
const Tables<float>& tables_of<float>(const d4b200_tables* h) { return h->t32; }

template <typename T>
int param_vjp(d4b200_tables_t h, const d4b200_params* par, int nbatch, int nat, const int64_t* numbers,
              const T* pos, const T* c6q, const T* c60, const T* gin, double* out, void* ws, size_t ws_size,
              void* stream) {
  if (!h || !par || nbatch < 0 || nat < 0) return D4B200_EINVAL;
  if (nbatch == 0) return 0;
  if (!out || nbatch > 65535) return D4B200_EINVAL;
  if (!(par->a1 == par->a1) || !(par->a2 == par->a2)) return D4B200_EPARAM;
  cudaStream_t st = (cudaStream_t)stream;
  if (nat == 0) return (int)cudaMemsetAsync(out, 0, (size_t)nbatch * 7 * sizeof(double), st);
  if (!numbers || !pos || !c6q || !c60 || !ws) return D4B200_EINVAL;
  if (ws_size < ws_bytes(nbatch, nat)) return D4B200_EWORKSPACE;
  ParV P;
  P.s6 = par->s6;
  P.s8 = par->s8;
  P.s9 = par->s9;
  P.has_s10 = par->has_s10 != 0;
  P.s10k = P.has_s10 ? par->s10 * 49.0 / 40.0 : 0.0;
  P.a1 = par->a1;
  P.a2 = par->a2;
  P.alp3 = par->alp / 3.0;
  P.disp2_sq = par->disp2_cutoff * par->disp2_cutoff;
  P.disp3_sq = par->disp3_cutoff * par->disp3_cutoff;
  double* planes = static_cast<double*>(ws);
  double* part2 = planes + (size_t)NPLANE * nbatch * nat * nat;
  double* part3 = part2 + (size_t)N2 * nbatch * nat;
  const dim3 grid((unsigned)nat, (unsigned)nbatch);
  k_param_pairs<T><<<grid, 128, 0, st>>>(tables_of<T>(h), P, nat, numbers, pos, c6q, c60, gin, planes, part2);
  k_param_triples<T><<<grid, 256, 0, st>>>(P, nat, numbers, gin, planes, part3);
  k_param_reduce<<<(unsigned)nbatch, 32, 0, st>>>(nat, part2, part3, out);
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" {

size_t d4b200_param_vjp_workspace_bytes(int nbatch, int nat) {
  if (nbatch <= 0 || nat <= 0) return 0;
  return ws_bytes(nbatch, nat);
}
int d4b200_param_vjp_f64(d4b200_tables_t h, const d4b200_params* par, int nbatch, int nat,
                         const int64_t* numbers_dev, const double* positions_dev, const double* c6q_dev,
                         const double* c60_dev, const double* grad_energy_dev, double* out_dev,
                         void* workspace_dev, size_t workspace_bytes, void* stream) {
  return param_vjp<double>(h, par, nbatch, nat, numbers_dev, positions_dev, c6q_dev, c60_dev, grad_energy_dev,
                           out_dev, workspace_dev, workspace_bytes, stream);
}
int d4b200_param_vjp_f32(d4b200_tables_t h, const d4b200_params* par, int nbatch, int nat,
                         const int64_t* numbers_dev, const float* positions_dev, const float* c6q_dev,
                         const float* c60_dev, const float* grad_energy_dev, double* out_dev,
                         void* workspace_dev, size_t workspace_bytes, void* stream) {
  return param_vjp<float>(h, par, nbatch, nat, numbers_dev, positions_dev, c6q_dev, c60_dev, grad_energy_dev,
                          out_dev, workspace_dev, workspace_bytes, stream);
}

}  // extern "C"

// One-CTA-per-structure kernels for molecules with <= 128 atoms ("small family").
//
// A persistent CTA pulls structures of one size class from a device-side queue
// (sorted by descending atom count), stages the atoms (SoA) and all per-pair
// quantities of the structure in shared memory and runs the whole D4 pipeline
// without touching HBM again:
//
//   compact -> CN (erfc count) -> Gaussian weights x zeta -> per-atom weighted
//   polarizability vectors A_i[23] -> pair pass (C6 = A_i.A_j, BJ two-body)
//   -> ATM pair stash (r^2, sigma/r^5, (R0/r)^(alp/3)) -> triple loop
//   [-> back-propagation passes for the gradient kernel]
//
// The kernel is templated on the class capacity CAP so that every shared-memory
// plane offset is an immediate.  Algebra and its derivation:
// tests/kernel_model.py (checked against the oracle on the CPU); reference
// formulation: src/tad_dftd4/dispersion/{twobody,threebody}.py, model/d4.py,
// tad_mctc.ncoord.cn_d4 (see include/d4b200.h).
#pragma once

#include <math.h>

#include "d4b200_common.cuh"
#include "d4b200_small_args.cuh"

namespace d4b200 {


// ---------------------------------------------------------------- bulk-async (TMA) staging
// The inputs of a structure are three contiguous rows of the padded batch (numbers, positions,
// charges).  A persistent CTA claims its NEXT structure one iteration early and has the TMA unit
// copy the three rows into a region of shared memory that is dead at that point
// (cp.async.bulk.shared::cluster.global, completion on an mbarrier), so that the compaction phase of
// the next iteration reads shared memory instead of waiting on global loads.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ unsigned r16(size_t b) { return (unsigned)((b + 15) & ~size_t(15)); }

// ---------------------------------------------------------------- math shims
__device__ __forceinline__ double d4_erfc(double x) { return erfc(x); }
__device__ __forceinline__ float d4_erfc(float x) { return erfcf(x); }
__device__ __forceinline__ double d4_exp(double x) { return exp(x); }
__device__ __forceinline__ float d4_exp(float x) { return expf(x); }
__device__ __forceinline__ double d4_log(double x) { return log(x); }
__device__ __forceinline__ float d4_log(float x) { return logf(x); }
__device__ __forceinline__ double d4_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float d4_sqrt(float x) { return sqrtf(x); }
// 1/sqrt(x) for normal positive x without the special-case branches of sqrt()/division:
// MUFU.RSQ64H seed (~20 bits) + one cubically convergent step y (1 + h/2 + 3 h^2/8),
// h = 1 - x y^2 (relative error ~ h^3 < 1e-17).  sqrt(x) = x rsqrt(x).
__device__ __forceinline__ double d4_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = fma(-(x * y), y, 1.0);
  return fma(y * h, fma(0.375, h, 0.5), y);
}
__device__ __forceinline__ float d4_rsqrt(float x) { return rsqrtf(x); }
// x^(-1/3) for normal positive x: float seed 2^(-log2(x)/3) from the MUFU units, one cubically
// convergent step z (1 + r/3 + 2 r^2/9), r = 1 - x z^3 (seed error ~1e-6 -> ~1e-17)
__device__ __forceinline__ double d4_rcbrt(double x) {
  const double z = (double)exp2f(-0.33333334f * __log2f((float)x));
  const double r = fma(-x, z * z * z, 1.0);
  return fma(z * r, fma(0.2222222222222222, r, 0.3333333333333333), z);
}
__device__ __forceinline__ float d4_rcbrt(float x) { return rcbrtf(x); }
// erfc(x) is below one ulp of the coordination number beyond this argument
__device__ __forceinline__ double d4_erfc_cut(double) { return 6.2; }
__device__ __forceinline__ float d4_erfc_cut(float) { return 4.6f; }

// Reciprocal without the slow-path branch of IEEE division: all operands here are normal,
// positive numbers.  MUFU.RCP64H seed (~20 bits) + one cubically convergent step
// y (1 + e + e^2), e = 1 - x y: three FMAs, relative error ~ e^3 < 1e-17.
__device__ __forceinline__ double d4_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  return fma(y, fma(e, e, e), y);
}
// float: MUFU.RCP seed (~1 ulp) + one Newton step; __frcp_rn is a ~16-instruction IEEE-rounding
// sequence that made up a third of the instructions of the FP32 gradient kernel (ncu source view)
__device__ __forceinline__ float d4_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return fmaf(y, fmaf(-x, y, 1.0f), y);
}
// Damping reciprocal of the gradient triple sweep: MUFU seed + ONE Newton step (two FMAs).
// Relative error ~ e^2 < 1e-12 on the three-body terms only, which are ~1e-2 of the energy
// and ~1e-5 Eh/Bohr in the gradient: four orders of magnitude inside the 1e-10 / 1e-9 bars.
__device__ __forceinline__ double d4_rcp_sweep(double x) {
#ifdef D4_RCP_SWEEP_CUBIC
  return d4_rcp(x);
#else
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return fma(y, fma(-x, y, 1.0), y);
#endif
}
__device__ __forceinline__ float d4_rcp_sweep(float x) {  // three-body terms only: the MUFU result as it is
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T>
__device__ __forceinline__ T d4_eps();
template <>
__device__ __forceinline__ double d4_eps<double>() { return 2.220446049250313e-16; }
template <>
__device__ __forceinline__ float d4_eps<float>() { return 1.1920929e-07f; }

// (R0/r)^(alp/3): x^6 (x^(-1/3))^2 for the default alp = 16, exp(y log x) otherwise
template <typename T>
__device__ __forceinline__ T d4_zero_damp_arg(T x, T alp3, bool alp16) {
  if (alp16) {
    const T z = d4_rcbrt(x);
    const T x3 = x * x * x;
    return (x3 * x3) * (z * z);
  }
  return d4_exp(alp3 * d4_log(x));
}

// stash entries of one pair from r^2, |C6(q=0)| and R0: 1/r, P' = fac9 sqrt|C6| / r^5
template <typename T>
__device__ __forceinline__ T d4_stash_p(T fac9, T c60, T rinv) {
  const T ac = fabs(c60);
  const T ri2 = rinv * rinv;
  const T sq = ac > T(1e-30) ? ac * d4_rsqrt(ac) : T(0);
  return fac9 * sq * (ri2 * ri2 * rinv);
}

// ---------------------------------------------------------------- smem layout
constexpr size_t al16(size_t b) { return (b + 15) & ~size_t(15); }

// per-atom T arrays
// (the D4 energy kernel only keeps the first five)
enum { AT_X = 0, AT_Y, AT_Z, AT_RCOV, AT_SQ, AT_Q, AT_CN, AT_E, AT_G, AT_DCN, AT_DQ };
// per-atom x 7 T arrays
enum { WT_Q = 0, WT_0, WT_ZGD, WT_Z0GD, WT_DZG };

// Row stride of the A/B vector buffers [23][AS]: the smallest value >= CAP that is
// 4 (mod 16), which makes the 8-atom x 4-frequency operand fragments of the FP64
// tensor-core contraction (mma.m8n8k4) bank-conflict free.
constexpr int a_stride(int cap) { return cap + (4 - cap % 16 + 16) % 16; }

// D4S: the Gaussian weights of an atom depend on the partner only through the partner's ELEMENT
// (model/d4s.py:61-67, 191-198), so they are tabulated once per (atom, distinct element of the
// structure) instead of being re-evaluated for every pair (14 float64 exponentials per side).
// Structures with more than D4S_ECAP distinct elements use the per-pair evaluation
// (D4S_ECAP, D4S_WSTR: d4b200_common.cuh).

template <typename T, bool GRAD, bool D4S, int CAP>
struct Lay {
  static constexpr int CP = CAP * (CAP - 1) / 2;
  static constexpr int AS = a_stride(CAP);
  static constexpr size_t plane_bytes = size_t(3) * CP * sizeof(T);
  // D4 energy kernel: the weights never leave the registers of the fused weights phase
  static constexpr bool wt_alias = false;
  // A/B vector buffers [23][AS]: D4 energy 2 (Aq, A0), D4 gradient 4 (Aq, A0, Bq, B0);
  // D4S has no per-atom vectors (pair-dependent weights) and only needs room for the
  // per-warp partial sums of the energy triple loop (16 warps)
  static constexpr size_t abuf_elems = D4S ? (GRAD ? 0 : size_t(16) * CAP) : size_t(GRAD ? 4 : 2) * NFREQ * AS;
  static constexpr int n_atom = GRAD ? 11 : (D4S ? 8 : 5);
  static constexpr int n_wt = D4S ? (GRAD ? 3 : 2) : (GRAD ? 5 : 0);
  static constexpr size_t planes = 0;
  static constexpr size_t abuf = al16(plane_bytes);
  static constexpr size_t atoms = abuf + al16(abuf_elems * sizeof(T));
  static constexpr size_t wts = atoms + al16(size_t(n_atom) * CAP * sizeof(T));
  static constexpr size_t ints = wts + al16(size_t(n_wt) * NREF * CAP * sizeof(T));
  // zs[CAP], idx[CAP], (D4S: element index per atom [CAP]), misc[32]
  // D4S gradient kernels: the reference-C6 blocks rc6[Z, Z', 7, 7] of the structure's distinct element
  // pairs, staged once per structure (the per-pair 7 x 7 contractions otherwise gather 49 values per
  // pair from the global table with lane-varying (Z, Z'): ncu shows them stalled on those loads)
  static constexpr size_t rs_elems = (D4S && GRAD) ? size_t(D4S_RCAP) * D4S_RCAP * NREF * NREF : 0;
  static constexpr size_t rs = ints + al16(((D4S ? 3 : 2) * CAP + 32) * sizeof(int));
  static constexpr size_t total = rs + al16(rs_elems * sizeof(T));
  static constexpr int scratch_planes = GRAD ? 5 : 3;  // gradient: Gamma, D, E3 shares (2), E2; energy: E3 shares (2), E2
  // per-CTA stride of the L2 scratch: the planes, then (D4S) the weight table [CAP][D4S_ECAP][D4S_WSTR]
  static constexpr size_t scratch_stride = size_t(scratch_planes) * CP + (D4S ? size_t(CAP) * D4S_ECAP * D4S_WSTR : 0);
};

// Unnormalised Gaussian weights S_a (and dS_a/dcn) of one atom for a given weighting
// factor, float64, max-shifted exponentials (model/d4s.py:189-213: pair-dependent wf).
template <bool DERIV>
__device__ __forceinline__ void d4s_gauss(const double* __restrict__ refcn, const int* __restrict__ refc,
                                          int z, double cn, double wf, double (&S)[NREF],
                                          double (&dS)[NREF], double& norm, double& dnorm) {
  double arg[NREF];
  double shift = 1e300;
#pragma unroll
  for (int a = 0; a < NREF; ++a) {
    const double d = cn - refcn[z * NREF + a];
    arg[a] = refc[z * NREF + a] > 0 ? wf * d * d : 1e300;
    shift = fmin(shift, arg[a]);
  }
  norm = 0.0;
  dnorm = 0.0;
#pragma unroll
  for (int a = 0; a < NREF; ++a) {
    const int rc = refc[z * NREF + a];
    const double d = cn - refcn[z * NREF + a];
    double s = 0.0, ds = 0.0;
    if (rc > 0) {  // geometric series in x = exp(-arg), see the weights phase of the kernel
      const double t1 = exp(shift - arg[a]), x = exp(-arg[a]);
      double pw = 1.0, acc = 0.0, dacc = 0.0;
      for (int k = 1; k <= rc; ++k) {
        acc += pw;
        dacc += (double)k * pw;
        pw *= x;
      }
      s = t1 * acc;
      if (DERIV) ds = -2.0 * wf * d * t1 * dacc;
    }
    S[a] = s;
    dS[a] = ds;
    norm += s;
    dnorm += ds;
  }
}

// normalised weights gw_a (and d gw_a/d cn) of one atom for a weighting factor, as the compute type
template <typename T, bool DERIV>
__device__ __forceinline__ void d4s_weights(const double* __restrict__ refcn, const int* __restrict__ refc, int z,
                                            double cn, double wf, T (&g)[NREF], T (&dg)[NREF]) {
  double S[NREF], dS[NREF], norm, dnorm;
  d4s_gauss<DERIV>(refcn, refc, z, cn, wf, S, dS, norm, dnorm);
  const double inv = norm > 0.0 ? 1.0 / norm : 0.0;
#pragma unroll
  for (int a = 0; a < NREF; ++a) {
    const double gg = S[a] * inv;
    g[a] = (T)gg;
    dg[a] = DERIV ? (T)((dS[a] - gg * dnorm) * inv) : T(0);
  }
}

// p -> (hi, lo) with hi > lo and p = hi(hi-1)/2 + lo: one L1-resident table lookup
// (Tables::pij) instead of a float square root plus corrections in every pair pass
__device__ __forceinline__ void pair_lookup(const unsigned short* __restrict__ pij, int p, int& hi, int& lo) {
  unsigned short v16;  // keep the table resident in the (small) L1 next to the streaming traffic
  asm("ld.global.nc.L1::evict_last.u16 %0, [%1];" : "=h"(v16) : "l"(pij + p));
  const unsigned v = v16;
  hi = (int)(v >> 8);
  lo = (int)(v & 255u);
}

// A_i . A_j over the 23 frequencies with four independent accumulators (a single
// chain of 23 dependent FMAs is latency bound)
template <typename T, int AS>
__device__ __forceinline__ T dot23(const T* __restrict__ A, int i, int j) {
  T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
#pragma unroll
  for (int w = 0; w + 4 <= NFREQ; w += 4) {
    s0 += A[w * AS + i] * A[w * AS + j];
    s1 += A[(w + 1) * AS + i] * A[(w + 1) * AS + j];
    s2 += A[(w + 2) * AS + i] * A[(w + 2) * AS + j];
    s3 += A[(w + 3) * AS + i] * A[(w + 3) * AS + j];
  }
  s0 += A[20 * AS + i] * A[20 * AS + j];
  s1 += A[21 * AS + i] * A[21 * AS + j];
  s2 += A[22 * AS + i] * A[22 * AS + j];
  return (s0 + s1) + (s2 + s3);
}

// C6_ij = A_i . A_j for all pairs of a structure on the FP64 tensor path: a warp owns
// an 8 x 8 tile of the (lower triangle of the) pair matrix and contracts the 23 (padded
// to 24) frequencies in six mma.m8n8k4 steps; results go to the packed pair plane
// `out[i (i-1)/2 + j]`, j < i.  Called by whole warps.
template <int AS>
__device__ __forceinline__ void c6_tiles(const double* __restrict__ Av, double* __restrict__ out,
                                         const unsigned short* __restrict__ pij, int n, int warp,
                                         int lane, int nwarps) {
  const int nb = (n + 7) >> 3;
  const int r = lane >> 2, c = lane & 3;
  for (int tile = warp; tile < nb * (nb + 1) / 2; tile += nwarps) {
    int I, J;  // tile = I (I + 1) / 2 + J, J <= I
    pair_lookup(pij, tile, I, J);
    I -= 1;
    const int i = I * 8 + r, j = J * 8 + r;
    double d0 = 0.0, d1 = 0.0;
#pragma unroll
    for (int k0 = 0; k0 < 24; k0 += 4) {
      const int w = k0 + c;
      const bool wok = w < NFREQ;
      const double av = wok ? Av[w * AS + i] : 0.0;
      const double bv = wok ? Av[w * AS + j] : 0.0;
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(d0), "+d"(d1)
                   : "d"(av), "d"(bv));
    }
    // accumulator fragment: row r (atom i), columns 2c, 2c+1 (atoms J*8 + ...)
    if (i < n) {
      const int jj = J * 8 + 2 * c, ti = i * (i - 1) / 2;
      if (jj < i) out[ti + jj] = d0;
      if (jj + 1 < i) out[ti + jj + 1] = d1;
    }
  }
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum_j lo[pair(i,j)] over j < i  +  sum_j hi[pair(j,i)] over j > i (the row phases are
// latency bound: eight lanes share a row)
// Lane `sub` of the row takes every eighth column; must be called by all 32 lanes of a warp, rows >= n contribute nothing.
template <typename T>
__device__ __forceinline__ T row_sum2_8(const T* __restrict__ lo, const T* __restrict__ hi, int i, int sub, int n) {
  T s = T(0);
  if (i < n) {
    const T* r = lo + i * (i - 1) / 2;
    for (int j = sub; j < i; j += 8) s += r[j];
    for (int j = i + 1 + sub; j < n; j += 8) s += hi[j * (j - 1) / 2 + i];
  }
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  return s;
}
#define D4_ROWS8(i, sub) \
  for (int t0_ = warp * 32, i = (t0_ + lane) >> 3, sub = lane & 7; t0_ < 8 * n; t0_ += NT, i = (t0_ + lane) >> 3)

// Gradient triple visit: owner pair (j,k) with r^2 = b, third atom i with the
// stash entries of (i,j) and (i,k).
//   e' = (0.375 s + abc) * P'_ij P'_ik P'_jk / (1 + 6 u_ij u_ik u_jk),  P' = P / r^2
// (threebody.py:113-160 factorised per pair; e' = e_ijk/6 with s9 folded in; the stash
// holds P' so that 1/(abc) never has to be formed).  With f = 1/(1 + 6 t):
//   d e'/d b = [ e' (alp f t - 2.5) + pf abc ] / b + 0.375 pf ds/db,   f t = (1 - f)/6
// The sweep accumulates  G = sum W e',  C = sum W [..],  S = sum W pf ds/db  and the caller
// forms  D = C / b + 0.375 S  once per owner pair (1/b and 0.375 are loop invariants).
// UNIT (closed structure, no upstream weights: d(sum E)): W = 6 for every triple, the sums
// are accumulated unweighted and G doubles as the energy share.  21 FP64 operations per
// visit (UNIT), one MUFU.
template <typename T, bool OPEN, bool UNIT>
__device__ __forceinline__ void grad_visit(T a_s, T Pij, T uij, T c_s, T Pik, T uik, T b, T b2, T twob,
                                           T cjk, T iP, T sPu, T kAi, T kB, T gi2, T gjk2,
                                           T gj, T gk, T& accG, T& accC, T& accS, T& accH, T& accL) {
  T a = a_s, c = c_s;
  T W = gi2 + gjk2;  // closed triple: multiplicity 2 for every atom (g*2 = 2 g*)
  T mj = T(2), mk = T(2);
  if (OPEN) {
    const T cij = a_s > T(0) ? T(1) : T(0);
    const T cik = c_s > T(0) ? T(1) : T(0);
    a = fabs(a_s);
    c = fabs(c_s);
    mj = cik * (cij + cjk);
    mk = cij * (cik + cjk);
    W = T(0.5) * gi2 * (cjk * (cij + cik)) + gj * mj + gk * mk;
  }
  // X = a+b-c, Y = a-b+c, Z = b+c-a:  X Z = b^2 - (a-c)^2,  X + Z = 2b
  const T t1 = a - c, Y = (a + c) - b;
  const T XZ = fma(-t1, t1, b2);
  const T s = XZ * Y;
  const T dsdb = fma(twob, Y, -XZ);  // d s / d b = Y Z - X Z + X Y
  const T abc = (a * c) * b;
  // owner-pair invariants folded into the damping denominator: iP = 1/P'_jk, sPu = 6 u_jk / P'_jk,
  // so that fp = P'_jk f = 1 / (iP + sPu u_ij u_ik) and pf = P'_ij P'_ik fp
  const T fp = d4_rcp_sweep(fma(sPu, uij * uik, iP));
  const T pf = (Pij * Pik) * fp;  // P' = P / r^2: P_ij P_ik P_jk / (abc d)
  // e' = pf wa;  e' (alp f t - 2.5) + pf abc = pf (wa k + abc) with
  // k = alp f t - 2.5 = (alp/6 - 2.5) - (alp/6) f,  f = fp / P'_jk  (kAi = alp / (6 P'_jk)):
  // pf is applied in the accumulating FMAs, e' itself is never formed
  const T wa = fma(T(0.375), s, abc);
  const T inner = fma(wa, fma(-kAi, fp, kB), abc);
  if (UNIT && !OPEN) {
    accG = fma(pf, wa, accG);
    accC = fma(pf, inner, accC);
    accS = fma(pf, dsdb, accS);
  } else {
    const T Wpf = W * pf;
    accG = fma(Wpf, wa, accG);
    accC = fma(Wpf, inner, accC);
    accS = fma(Wpf, dsdb, accS);
    if (OPEN) {
      const T e = pf * wa;
      accH = fma(mj, e, accH);  // energy shares of the owner pair's atoms (fused energy + gradient call)
      accL = fma(mk, e, accL);
    } else {
      accH = fma(pf, wa, accH);  // both atoms of the owner pair have multiplicity 2 (applied by the caller)
    }
  }
}

// One block of 8 consecutive top atoms i0..i0+7 for the lane's bottom pair (j,k).
// PRED = false: every lane is active for all eight rows -> straight-line code in which
// the compiler interleaves the eight independent evaluations (the chain of one
// evaluation is ~14 dependent FP64 operations); PRED = true: per-row predicates.
// Only three instantiations exist (closed straight, closed predicated, open
// predicated) so that the hot loop stays inside the instruction cache.
template <typename T, int CP, bool PRED, bool OPEN>
__device__ __forceinline__ void triple_block8(const T* __restrict__ colj, const T* __restrict__ colk,
                                              int i0, int j, int n, T bb, T bb2, T cjk, T Pjk,
                                              T ujk, T& accJ, T& accK, T (&v)[8]) {
  int ti = i0 * (i0 - 1) / 2;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int i = i0 + u;
    T ei = T(0);
    if (!PRED || (i < n && i > j)) {
      const T a_s = colj[ti], c_s = colk[ti];
      const T t = colj[ti + 2 * CP] * colk[ti + 2 * CP] * ujk;
      const T pp = colj[ti + CP] * colk[ti + CP] * Pjk;
      T a = a_s, c = c_s;
      T mi = T(1), mj = T(1), mk = T(1);
      if (OPEN) {
        const T cij = a_s > T(0) ? T(1) : T(0);
        const T cik = c_s > T(0) ? T(1) : T(0);
        a = fabs(a_s);
        c = fabs(c_s);
        mi = cjk * (cij + cik);
        mj = cik * (cij + cjk);
        mk = cij * (cik + cjk);
      }
      // s = (a+b-c)(a-b+c)(b+c-a) = (b^2 - (a-c)^2) (a+c-b);  e = P'P'P' (0.375 s + abc) / d
      const T t1 = a - c, t2 = a + c;
      const T s = fma(-t1, t1, bb2) * (t2 - bb);
      const T abc = (a * c) * bb;
      const T e = (pp * fma(T(0.375), s, abc)) * d4_rcp(fma(T(6), t, T(1)));
      if (OPEN) {
        accJ += mj * e;
        accK += mk * e;
        ei = mi * e;
      } else {
        ei = e;
      }
    }
    v[u] = ei;
    ti += i;
  }
  // closed triples: the share of the bottom pair is the plain sum of the block (tree sum
  // instead of a chain of eight dependent additions)
  if (!OPEN) accJ += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
}

// transposed reduction: 8 values x 32 lanes -> lane group (lane>>2) holds the sum of v[lane>>2]
template <typename T>
__device__ __forceinline__ T reduce8x32(const T (&v)[8], bool b4, bool b3, bool b2) {
  const T s0 = b4 ? v[0] : v[4], s1 = b4 ? v[1] : v[5], s2 = b4 ? v[2] : v[6], s3 = b4 ? v[3] : v[7];
  const T w0 = (b4 ? v[4] : v[0]) + __shfl_xor_sync(0xffffffffu, s0, 16);
  const T w1 = (b4 ? v[5] : v[1]) + __shfl_xor_sync(0xffffffffu, s1, 16);
  const T w2 = (b4 ? v[6] : v[2]) + __shfl_xor_sync(0xffffffffu, s2, 16);
  const T w3 = (b4 ? v[7] : v[3]) + __shfl_xor_sync(0xffffffffu, s3, 16);
  const T x0 = (b3 ? w2 : w0) + __shfl_xor_sync(0xffffffffu, b3 ? w0 : w2, 8);
  const T x1 = (b3 ? w3 : w1) + __shfl_xor_sync(0xffffffffu, b3 ? w1 : w3, 8);
  T y = (b2 ? x1 : x0) + __shfl_xor_sync(0xffffffffu, b2 ? x0 : x1, 4);
  y += __shfl_xor_sync(0xffffffffu, y, 2);
  y += __shfl_xor_sync(0xffffffffu, y, 1);
  return y;
}

// ---------------------------------------------------------------------------
// Register-tiled gradient sweep (closed structure, unit upstream weights).
//
// The owner-pair sweep above reads six stash entries (48 B) per visit for 20 FP64
// instructions: four warps per scheduler at the FP64 rate ask the shared-memory
// crossbar (128 B/clk/SM) for ~0.9 wavefronts per clock -- it is co-limited by LDS
// bandwidth.  Here a thread owns the 2 x 2 block of owner pairs {i0, i1} x {k0, k1},
//   i0 = 2I, i1 = 2I + 1, k0 = m, k1 = m + I   (0 <= m < I),
// and sweeps the third atom j: the stash entries of (i0,j), (i1,j), (k0,j), (k1,j) are
// loaded once and serve four visits (24 B per visit).  Consecutive lanes take
// consecutive m, so the k loads are unit-stride across the warp (conflict free) and the
// i loads are broadcasts.  The blocks cover every pair (i,k) with k < 2 (i >> 1) exactly
// once; the remaining "diagonal" pairs (2I + 1, 2I) are swept by eight lanes each.
// When the last round of blocks fills less than half of the CTA, every block is shared
// by 2 or 4 lanes (sub-ranges of j, combined with shuffles in a fixed order).
// ---------------------------------------------------------------------------

// per-pair epilogue of the closed/unit sweep: sums over the third atom -> dL/dC6(q=0), dL/dr^2
// and the energy shares of the fused energy + gradient call
template <typename T, bool D4S, int CP>
__device__ __forceinline__ void grad_pair_store(T* __restrict__ out0, T* __restrict__ out1, bool want_e, int p,
                                                T bb, T Pjk, T ifac9, T accG, T accC, T accS) {
  const bool pz = Pjk == T(0);  // no ATM contribution through this pair (C6(q=0) = 0)
  T accH = accG + accG;         // both atoms of the owner pair have multiplicity 2
  accG *= T(6);                 // unit upstream weights: W = 6 for every triple
  accC *= T(6);
  accS *= T(6);
  if (pz) accG = accC = accS = accH = T(0);
  const T accD = fma(accC, d4_rcp(bb), T(0.375) * accS);
  if constexpr (!D4S) {
    const T w = (Pjk * ifac9) * (bb * bb);  // C6(q=0) = (P'_jk r^5 / fac9)^2
    accG = pz ? T(0) : accG * (T(0.5) * d4_rcp((w * w) * bb));
  }
  out0[p] = accG;
  out1[p] = accD;
  if (want_e) {
    out0[2 * CP + p] = accH;
    out0[3 * CP + p] = accH;
  }
}

// loop invariants of one owner pair (see grad_visit)
template <typename T>
struct OwnerPair {
  T b, b2, twob, iP, sPu, kAi;
  __device__ __forceinline__ void load(const T* __restrict__ pa, const T* __restrict__ pP,
                                       const T* __restrict__ pu, int p, T kA) {
    b = fabs(pa[p]);
    b2 = b * b;
    twob = b + b;
    const T Pjk = pP[p];
    iP = Pjk == T(0) ? T(1) : d4_rcp(Pjk);
    sPu = T(6) * pu[p] * iP;
    kAi = kA * iP;
  }
};

template <typename T, bool D4S, int CAP, int NT>
__device__ __forceinline__ void grad_sweep_tiled(const T* __restrict__ pa, const T* __restrict__ pP,
                                                 const T* __restrict__ pu, T* __restrict__ out0,
                                                 T* __restrict__ out1, bool want_e,
                                                 const unsigned short* __restrict__ pij, int n, int tid,
                                                 T kA, T kB, T ifac9) {
  constexpr int CP = CAP * (CAP - 1) / 2;
  const int lane = tid & 31, warp = tid >> 5;
  const int nI = (n + 1) >> 1;  // atom pairs (2I, 2I + 1); the last one may lack its second atom
  const int nblk = nI * (nI - 1) / 2;
  const T z = T(0);
  for (int base = 0; base < nblk; base += NT) {
    const int left = nblk - base;
    // lanes per block in this round (1, 2 or 4; CTA-uniform)
    const int parts = left * 2 > NT ? 1 : (left * 4 > NT ? 2 : 4);
    const int per = 32 / parts;  // blocks per warp
    if (warp * per >= left) continue;
    const int sub = lane / per;  // which part of the j range
    const int item = base + warp * per + (lane - sub * per);
    const bool valid = item < nblk;
    int I, m;
    pair_lookup(pij, valid ? item : nblk - 1, I, m);
    const int i0 = 2 * I, k0 = m, k1 = m + I;
    const bool has1 = i0 + 1 < n;
    const int i1 = has1 ? i0 + 1 : i0;  // odd n: the last row pair is a single row (evaluated twice, stored once)
    const int ti0 = i0 * (i0 - 1) / 2, ti1 = i1 * (i1 - 1) / 2;
    const int tk0 = k0 * (k0 - 1) / 2, tk1 = k1 * (k1 - 1) / 2;
    const int p00 = ti0 + k0, p01 = ti0 + k1, p10 = ti1 + k0, p11 = ti1 + k1;
    OwnerPair<T> o00, o01, o10, o11;
    o00.load(pa, pP, pu, p00, kA);
    o01.load(pa, pP, pu, p01, kA);
    o10.load(pa, pP, pu, p10, kA);
    o11.load(pa, pP, pu, p11, kA);
    T g00 = z, c00 = z, s00 = z, g01 = z, c01 = z, s01 = z;
    T g10 = z, c10 = z, s10 = z, g11 = z, c11 = z, s11 = z;
    T dH = z, dL = z;  // unused outputs of grad_visit in this mode
    const int len = (n + parts - 1) / parts;
    const int jbeg = sub * len;
    const int jend = min(n, jbeg + len);
#ifndef D4_TILED_UNROLL
#define D4_TILED_UNROLL 1
#endif
    constexpr int kUnroll = D4_TILED_UNROLL;  // j steps in flight (four visits each)
#pragma unroll kUnroll
    for (int j = jbeg; j < jend; ++j) {
      const int tj = j * (j - 1) / 2;
      // stash index of (x, j); for j == x the owner's own entry (finite values, factor zeroed below)
      const int xi0 = j < i0 ? ti0 + j : (j > i0 ? tj + i0 : p00);
      const int xi1 = j < i1 ? ti1 + j : (j > i1 ? tj + i1 : p10);
      const int xk0 = j < k0 ? tk0 + j : (j > k0 ? tj + k0 : p00);
      const int xk1 = j < k1 ? tk1 + j : (j > k1 ? tj + k1 : p01);
      const T ai0 = pa[xi0], ui0 = pu[xi0], Pi0 = j != i0 ? pP[xi0] : z;
      const T ai1 = pa[xi1], ui1 = pu[xi1], Pi1 = j != i1 ? pP[xi1] : z;
      const T ak0 = pa[xk0], uk0 = pu[xk0], Pk0 = j != k0 ? pP[xk0] : z;
      const T ak1 = pa[xk1], uk1 = pu[xk1], Pk1 = j != k1 ? pP[xk1] : z;
      grad_visit<T, false, true>(ai0, Pi0, ui0, ak0, Pk0, uk0, o00.b, o00.b2, o00.twob, z, o00.iP, o00.sPu,
                                 o00.kAi, kB, z, z, z, z, g00, c00, s00, dH, dL);
      grad_visit<T, false, true>(ai0, Pi0, ui0, ak1, Pk1, uk1, o01.b, o01.b2, o01.twob, z, o01.iP, o01.sPu,
                                 o01.kAi, kB, z, z, z, z, g01, c01, s01, dH, dL);
      grad_visit<T, false, true>(ai1, Pi1, ui1, ak0, Pk0, uk0, o10.b, o10.b2, o10.twob, z, o10.iP, o10.sPu,
                                 o10.kAi, kB, z, z, z, z, g10, c10, s10, dH, dL);
      grad_visit<T, false, true>(ai1, Pi1, ui1, ak1, Pk1, uk1, o11.b, o11.b2, o11.twob, z, o11.iP, o11.sPu,
                                 o11.kAi, kB, z, z, z, z, g11, c11, s11, dH, dL);
    }
    if (parts > 1) {  // combine the sub-ranges: fixed order, the first lane of a block ends up with the sum
#define D4_COMBINE(v)                                                     \
  if (parts == 4) v += __shfl_down_sync(0xffffffffu, v, 16);              \
  v += __shfl_down_sync(0xffffffffu, v, per);
      // parts == 4: per = 8, lanes l, l+8, l+16, l+24 -> (l, l+16) and (l+8, l+24), then l + (l+8)
      D4_COMBINE(g00) D4_COMBINE(c00) D4_COMBINE(s00) D4_COMBINE(g01) D4_COMBINE(c01) D4_COMBINE(s01)
      D4_COMBINE(g10) D4_COMBINE(c10) D4_COMBINE(s10) D4_COMBINE(g11) D4_COMBINE(c11) D4_COMBINE(s11)
#undef D4_COMBINE
    }
    if (valid && sub == 0) {
      grad_pair_store<T, D4S, CP>(out0, out1, want_e, p00, o00.b, pP[p00], ifac9, g00, c00, s00);
      grad_pair_store<T, D4S, CP>(out0, out1, want_e, p01, o01.b, pP[p01], ifac9, g01, c01, s01);
      if (has1) {
        grad_pair_store<T, D4S, CP>(out0, out1, want_e, p10, o10.b, pP[p10], ifac9, g10, c10, s10);
        grad_pair_store<T, D4S, CP>(out0, out1, want_e, p11, o11.b, pP[p11], ifac9, g11, c11, s11);
      }
    }
  }
  // diagonal pairs (2d + 1, 2d): eight lanes per pair, each an eighth of the j range
  const int nd = n >> 1;
  const int len8 = (n + 7) >> 3;
  for (int d0 = 0; d0 < nd; d0 += NT / 8) {
    if (d0 + warp * 4 >= nd) continue;
    const int d = d0 + (tid >> 3), sub = lane & 7;
    const bool valid = d < nd;
    const int jj = 2 * (valid ? d : nd - 1) + 1, kk = jj - 1;
    const int tjj = jj * (jj - 1) / 2, tkk = kk * (kk - 1) / 2;
    const int p = tjj + kk;
    OwnerPair<T> o;
    o.load(pa, pP, pu, p, kA);
    T g = z, c = z, s = z, dH = z, dL = z;
    const int jend = min(n, (sub + 1) * len8);
#pragma unroll 2
    for (int j = sub * len8; j < jend; ++j) {
      const int tj = j * (j - 1) / 2;
      const bool ok = (j != jj) & (j != kk);
      const int x1 = ok ? (j < jj ? tjj + j : tj + jj) : p;
      const int x2 = ok ? (j < kk ? tkk + j : tj + kk) : p;
      grad_visit<T, false, true>(pa[x1], ok ? pP[x1] : z, pu[x1], pa[x2], pP[x2], pu[x2], o.b, o.b2, o.twob, z,
                                 o.iP, o.sPu, o.kAi, kB, z, z, z, z, g, c, s, dH, dL);
    }
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) {
      g += __shfl_down_sync(0xffffffffu, g, off);
      c += __shfl_down_sync(0xffffffffu, c, off);
      s += __shfl_down_sync(0xffffffffu, s, off);
    }
    if (valid && sub == 0) grad_pair_store<T, D4S, CP>(out0, out1, want_e, p, o.b, pP[p], ifac9, g, c, s);
  }
}

template <typename T, bool GRAD, bool D4S, int CAP, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) small_kernel(SmallArgs<T> A) {
  using L = Lay<T, GRAD, D4S, CAP>;
  constexpr int CP = L::CP;
  constexpr int AS = L::AS;
  constexpr int NW = NT / 32;
  // D4 energy kernel (FP64): once the tensor-path tiles have consumed the A vectors, the
  // buffer holds the two-body pair energies [CP] followed by the per-warp energy partials
  // [NW][CAP]; otherwise (too small a buffer, FP32) the pair energies go to the L2 scratch
  constexpr bool E2S = !GRAD && !D4S && sizeof(T) == 8 &&
                       size_t(CP) + size_t(NW) * CAP <= size_t(2) * NFREQ * AS;
  extern __shared__ __align__(16) unsigned char smem[];
  T* const pa = reinterpret_cast<T*>(smem + L::planes);
  T* const pP = pa + CP;
  T* const pu = pP + CP;
  T* const Aq = reinterpret_cast<T*>(smem + L::abuf);
  T* const A0 = Aq + NFREQ * AS;
  T* const Bq = Aq + 2 * NFREQ * AS;  // GRAD only
  T* const B0 = Aq + 3 * NFREQ * AS;  // GRAD only
  T* const at = reinterpret_cast<T*>(smem + L::atoms);
  T* const wt = L::wt_alias ? reinterpret_cast<T*>(smem + L::planes) + CP
                            : reinterpret_cast<T*>(smem + L::wts);
  int* const zs = reinterpret_cast<int*>(smem + L::ints);
  int* const idx = zs + CAP;
  int* const ei = idx + CAP;  // D4S: index of the atom's element among the distinct elements of the structure
  // [0]=work item, [1]=n, [2]=any_open, [3]=skip, [4]=chunk counter, [5]=near pairs,
  // D4S: [8..11]=element bit set, [12..12+D4S_ECAP)=distinct elements
  int* const misc = idx + (D4S ? 2 : 1) * CAP;
#define ATOM(k) (at + (k) * CAP)
#define WT(k) (wt + (k) * NREF * CAP)

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const Par<T>& P = A.par;
  const Tables<T>& tab = A.tab;
  const int range_begin = A.wk.class_range[2 * A.cls];
  const int range_end = A.wk.class_range[2 * A.cls + 1];
  // per-CTA, L2-resident scratch planes for per-pair results
  T* const out0 = A.scratch + (size_t)blockIdx.x * L::scratch_stride;
  T* const out1 = out0 + CP;
  T* const wtab = out0 + size_t(L::scratch_planes) * CP;  // D4S weight table of the current structure
  long long tlast = A.phase ? clock64() : 0;
#define PHASE(k)                                                        \
  if (A.phase && tid == 0) {                                            \
    const long long now = clock64();                                    \
    atomicAdd(&A.phase[k], (unsigned long long)(now - tlast));          \
    tlast = now;                                                        \
  }

  // ---- bulk-async staging of the next structure's input rows (see the helpers above) ----------
  // misc[24] = next work item (-1: not claimed yet), misc[25] = its rows are being staged,
  // misc[26] = the current item's rows are staged, misc[28..29] = the mbarrier
  unsigned long long* const bar = reinterpret_cast<unsigned long long*>(misc + 28);
  constexpr size_t abuf_used = GRAD ? 2 * NFREQ * AS : ((E2S ? CP : 0) + size_t(NW) * CAP + 3) / 4 * 4;
  unsigned char* const stage = reinterpret_cast<unsigned char*>(Aq + abuf_used);  // gradient: Bq, B0
  constexpr size_t stage_avail = D4S ? 0 : (size_t(GRAD ? 4 : 2) * NFREQ * AS - abuf_used) * sizeof(T);
  const unsigned seg_p = r16((size_t)A.nat * 8 + 16);
  const unsigned seg_q = seg_p + r16((size_t)A.nat * 3 * sizeof(T) + 16);
  const unsigned seg_end = seg_q + r16((size_t)A.nat * sizeof(T) + 16);
  // Gradient kernels only: in the energy kernels (80 registers per thread at three or more CTAs per SM)
  // the extra live pointers cost spills and 3 % of the C2 step, while their compaction phase is already
  // overlapped by the other resident CTAs (measured)
#ifdef D4_NO_STAGE  // A/B knob: plain global loads in the compaction phase
  constexpr bool STAGE_ON = false;
#else
  constexpr bool STAGE_ON = GRAD && !D4S;
#endif
  const bool use_stage = STAGE_ON && seg_end <= stage_avail;
  unsigned bar_parity = 0;
  if (tid == 0) {
    misc[24] = -1;
    misc[25] = 0;
    if (use_stage) mbar_init(bar, 1);
  }
  // claim the next work item and start the copy of its rows; called by thread 0 after a block barrier
  // behind which the staging region is dead
  auto prefetch_next = [&]() {
    const int nxt = atomicAdd(&A.wk.queue[A.cls], 1);
    int staged = 0;
    if (use_stage && range_begin + nxt < range_end) {
      const size_t o2 = (size_t)A.wk.order[range_begin + nxt] * A.nat, tot = (size_t)A.nbatch * A.nat;
      const char* z0 = reinterpret_cast<const char*>(A.numbers + o2);
      const char* p0 = reinterpret_cast<const char*>(A.pos + 3 * o2);
      const char* q0 = reinterpret_cast<const char*>(A.q + o2);
      const unsigned lz = (unsigned)((size_t)z0 & 15), lp = (unsigned)((size_t)p0 & 15), lq = (unsigned)((size_t)q0 & 15);
      const unsigned bz = r16(lz + (size_t)A.nat * 8), bp = r16(lp + (size_t)A.nat * 3 * sizeof(T)),
                     bq = r16(lq + (size_t)A.nat * sizeof(T));
      // the 16-byte aligned windows must stay inside the arrays (first / last rows of unaligned views)
      const bool inside = z0 - lz >= reinterpret_cast<const char*>(A.numbers) &&
                          z0 - lz + bz <= reinterpret_cast<const char*>(A.numbers + tot) &&
                          p0 - lp >= reinterpret_cast<const char*>(A.pos) &&
                          p0 - lp + bp <= reinterpret_cast<const char*>(A.pos + 3 * tot) &&
                          q0 - lq >= reinterpret_cast<const char*>(A.q) &&
                          q0 - lq + bq <= reinterpret_cast<const char*>(A.q + tot);
      if (inside) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy accesses of the region are done
        mbar_expect_tx(bar, bz + bp + bq);
        bulk_g2s(stage, z0 - lz, bz, bar);
        bulk_g2s(stage + seg_p, p0 - lp, bp, bar);
        bulk_g2s(stage + seg_q, q0 - lq, bq, bar);
        staged = 1;
      }
    }
    misc[24] = nxt;
    misc[25] = staged;
  };

  while (true) {
    __syncthreads();  // previous structure fully written, misc[] reusable
    if (tid == 0) {
      if (misc[24] < 0) {  // nothing claimed ahead (first iteration, properties mode)
        misc[24] = atomicAdd(&A.wk.queue[A.cls], 1);
        misc[25] = 0;
      }
      misc[0] = misc[24];
      misc[26] = misc[25];
      misc[24] = -1;
      misc[25] = 0;
    }
    __syncthreads();
    const int item = range_begin + misc[0];
    if (item >= range_end) break;
    const int b = A.wk.order[item];
    const bool staged = STAGE_ON && misc[26] != 0;
    if (staged) {  // the rows of this structure were copied into shared memory during the previous iteration
      mbar_wait(bar, bar_parity);
      bar_parity ^= 1u;
    }
    const int64_t* const zrow_g = A.numbers + (size_t)b * A.nat;
    const T* const prow_g = A.pos + (size_t)b * A.nat * 3;
    const T* const qrow_g = A.q + (size_t)b * A.nat;
    const int64_t* const zrow = staged ? reinterpret_cast<const int64_t*>(stage + ((size_t)zrow_g & 15)) : zrow_g;
    const T* const prow = staged ? reinterpret_cast<const T*>(stage + seg_p + ((size_t)prow_g & 15)) : prow_g;
    const T* const qrow = staged ? reinterpret_cast<const T*>(stage + seg_q + ((size_t)qrow_g & 15)) : qrow_g;

    // ---- phase 0: compact real atoms (numbers != 0), zero padded outputs ----
    if (tid < 32) {
      int count = 0, bad = 0;
      for (int base = 0; base < A.nat; base += 32) {
        const int src = base + lane;
        long long zv = src < A.nat ? zrow[src] : 0;
        const bool real = zv != 0;
        if (real && (zv < 0 || zv >= NELEM)) {
          bad = 1;
          zv = 1;
        }
        const unsigned m = __ballot_sync(0xffffffffu, real);
        const int dst = count + __popc(m & ((1u << lane) - 1u));
        if (real && dst < CAP) {
          idx[dst] = src;
          zs[dst] = (int)zv;
        }
        count += __popc(m);
      }
      bad = __any_sync(0xffffffffu, bad);
      if (lane == 0) {
        misc[1] = count <= CAP ? count : 0;
        misc[2] = 0;
        misc[3] = count > CAP;
        misc[4] = 0;
        misc[5] = 0;
        if (D4S) misc[8] = misc[9] = misc[10] = misc[11] = 0;
        if (bad) atomicOr(A.wk.status, D4B200_STATUS_BAD_NUMBER);
        if (count > CAP) atomicOr(A.wk.status, D4B200_STATUS_TOO_LARGE);
      }
    }
    __syncthreads();
    const bool skip = misc[3] != 0;
    for (int t = tid; t < A.nat; t += NT) {
      if (zrow[t] == 0 || skip) {
        const size_t o = (size_t)b * A.nat + t;
        if (!GRAD) {
          A.energy[o] = T(0);
          if (A.cn_out) A.cn_out[o] = T(0);
          if (A.alpha_out) A.alpha_out[o] = T(0);
        } else {
          if (A.energy) A.energy[o] = T(0);
          if (A.grad) {
            A.grad[3 * o] = T(0);
            A.grad[3 * o + 1] = T(0);
            A.grad[3 * o + 2] = T(0);
          }
          if (A.gradq) A.gradq[o] = T(0);
        }
      }
    }
    const int n = misc[1];
    const int np = n * (n - 1) / 2;

    for (int i = tid; i < n; i += NT) {
      const size_t o = (size_t)b * A.nat + idx[i];
      const int z = zs[i];
      ATOM(AT_X)[i] = prow[3 * idx[i]];
      ATOM(AT_Y)[i] = prow[3 * idx[i] + 1];
      ATOM(AT_Z)[i] = prow[3 * idx[i] + 2];
      if constexpr (GRAD || D4S) ATOM(AT_Q)[i] = qrow[idx[i]];
      ATOM(AT_RCOV)[i] = tab.rcov[z];
      ATOM(AT_SQ)[i] = tab.sqrt_r4r2[z];
      if (GRAD) ATOM(AT_G)[i] = A.gin ? A.gin[o] : T(1);
    }
    __syncthreads();
    PHASE(0);

    // ---- phase 1: coordination number (tad_mctc cn_d4 / erf_count) ----------
    // Sweep 1 (all pairs): squared distance, and a compacted list of the pairs whose
    // counting function is above one ulp (r < (1 + cut/kcn) r0, about five per atom).
    // Sweep 2 (listed pairs only): the erfc evaluation, on full warps.
    {
      int* const near = reinterpret_cast<int*>(pP);  // plane `pP` is unused until the weights
      const T reach = T(1) + d4_erfc_cut(T(0)) / T(7.5);
      for (int p0 = warp * 32; p0 < np; p0 += NT) {
        const int p = p0 + lane;
        bool hit = false;
        if (p < np) {
          int i, j;
          pair_lookup(tab.pij, p, i, j);
          const T dx = ATOM(AT_X)[i] - ATOM(AT_X)[j];
          const T dy = ATOM(AT_Y)[i] - ATOM(AT_Y)[j];
          const T dz = ATOM(AT_Z)[i] - ATOM(AT_Z)[j];
          const T r2 = dx * dx + dy * dy + dz * dz;
          pa[p] = r2;  // later pair passes read the squared distance from here
          pu[p] = T(0);
          const T rr = reach * (ATOM(AT_RCOV)[i] + ATOM(AT_RCOV)[j]);
          hit = r2 <= P.cn_sq && r2 < rr * rr;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) {
          int base = 0;
          if (lane == 0) base = atomicAdd(&misc[5], __popc(m));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (hit) near[base + __popc(m & ((1u << lane) - 1u))] = p;
        }
      }
      __syncthreads();
      const int nnear = misc[5];
      for (int t = tid; t < nnear; t += NT) {
        const int p = near[t];
        int i, j;
        pair_lookup(tab.pij, p, i, j);
        const T r2n = pa[p];
        const T r = r2n * d4_rsqrt(r2n);
        const T r0 = ATOM(AT_RCOV)[i] + ATOM(AT_RCOV)[j];
        const T xx = T(7.5) * (r * d4_rcp(r0) - T(1));
        if (xx < d4_erfc_cut(T(0))) pu[p] = tab.den[zs[i] * NELEM + zs[j]] * T(0.5) * d4_erfc(xx);
      }
    }
    __syncthreads();
    PHASE(1);
    // ---- phase 2: CN row sums -> Gaussian weights (float64 always) x zeta -> A_i[w] ----
    // model/d4.py:137-228; max-shifted exponentials instead of pow(exp(-d^2), k wf).
    // Eight lanes per atom (one per reference, one idle): the row sum, the shift (min
    // exponent) and the normalisation are 8-lane shuffle reductions; the weights stay in
    // registers for the weighted polarizability vectors (three frequencies per lane).
#ifndef D4_WEIGHTS_LANES8
    if constexpr (!D4S) {
      // D4 kernels: FOUR lanes per atom (references a and a + 4 per lane, frequencies
      // a, a + 4, ..., a + 20), so that the whole structure is one pass of the CTA for every
      // size class (CAP atoms x 4 lanes = NT threads): the phase is a single latency chain
      // (row sum -> shift -> exponentials -> normalisation -> vectors), two passes of eight
      // lanes per atom cost that chain twice.
      for (int t0 = warp * 32; t0 < 4 * n; t0 += NT) {
        const int i = (t0 + lane) >> 2, a = lane & 3;
        const bool row = i < n;
        const int z = row ? zs[i] : 0;
        // table entries first (independent of the row sum)
        int rc[2];
        double rcn[2], qref[2], z0[2];
        int za[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ar = a + 4 * h;
          za[h] = z * NREF + (ar < NREF ? ar : 0);
          rc[h] = (row && ar < NREF) ? tab.refc[za[h]] : 0;
          rcn[h] = tab.refcn[za[h]];
          qref[h] = tab.refq[za[h]];
          z0[h] = tab.zeta0[za[h]];
        }
        const double gam = tab.gamgc[z], zeff = tab.zeff[z];
        double qat = 0.0;
        if (row) {
          if constexpr (GRAD) qat = (double)ATOM(AT_Q)[i];
          else qat = (double)A.q[(size_t)b * A.nat + idx[i]];  // energy kernel: never staged
        }
        T c0 = T(0), c1 = T(0);
        if (row) {
          const T* r = pu + i * (i - 1) / 2;
          int j = a;
          for (; j + 4 < i; j += 8) {
            c0 += r[j];
            c1 += r[j + 4];
          }
          if (j < i) c0 += r[j];
          j = i + 1 + a;
          for (; j + 4 < n; j += 8) {
            c0 += pu[j * (j - 1) / 2 + i];
            c1 += pu[(j + 4) * (j + 3) / 2 + i];
          }
          if (j < n) c0 += pu[j * (j - 1) / 2 + i];
        }
        T cn_row = c0 + c1;
        cn_row += __shfl_xor_sync(0xffffffffu, cn_row, 2);
        cn_row += __shfl_xor_sync(0xffffffffu, cn_row, 1);
        if (a == 0 && row) {
          if constexpr (GRAD) ATOM(AT_CN)[i] = cn_row;
          else if (A.cn_out) A.cn_out[(size_t)b * A.nat + idx[i]] = cn_row;
        }
        double arg[2], dd[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const double d = (double)cn_row - rcn[h];
          dd[h] = d;
          arg[h] = rc[h] > 0 ? P.wf * d * d : 1e300;
        }
        double shift = fmin(arg[0], arg[1]);
        shift = fmin(shift, __shfl_xor_sync(0xffffffffu, shift, 2));
        shift = fmin(shift, __shfl_xor_sync(0xffffffffu, shift, 1));
        double S[2], dS[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const double t1 = exp(shift - arg[h]), x = exp(-arg[h]);
          double pw = 1.0, acc = 0.0, dacc = 0.0;
#pragma unroll
          for (int k = 1; k <= 3; ++k) {
            acc += k <= rc[h] ? pw : 0.0;
            if (GRAD) dacc += k <= rc[h] ? (double)k * pw : 0.0;
            pw *= x;
          }
          for (int k = 4; k <= rc[h]; ++k) {
            acc += pw;
            if (GRAD) dacc += (double)k * pw;
            pw *= x;
          }
          S[h] = rc[h] > 0 ? t1 * acc : 0.0;
          dS[h] = (GRAD && rc[h] > 0) ? -2.0 * P.wf * dd[h] * t1 * dacc : 0.0;
        }
        double norm = S[0] + S[1];
        norm += __shfl_xor_sync(0xffffffffu, norm, 2);
        norm += __shfl_xor_sync(0xffffffffu, norm, 1);
        double dnorm = 0.0;
        if constexpr (GRAD) {
          dnorm = dS[0] + dS[1];
          dnorm += __shfl_xor_sync(0xffffffffu, dnorm, 2);
          dnorm += __shfl_xor_sync(0xffffffffu, dnorm, 1);
        }
        const double qmod = qat + zeff;
        const bool qpos = qmod > 0.0;
        const double qinv = 1.0 / (qpos ? qmod - (double)d4_eps<T>() : 1.0);
        const bool nz = norm > 0.0;
        const double inv = d4_rcp(nz ? norm : 1.0);
        T wq[2], w0[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const double scale = qpos ? exp(gam * (1.0 - qref[h] * qinv)) : 0.0;
          const double zeta = rc[h] > 0 ? exp(P.ga * (1.0 - scale)) : 0.0;
          const double gw = nz ? S[h] * inv : 0.0;
          wq[h] = (T)(zeta * gw);
          w0[h] = (T)(z0[h] * gw);
          if constexpr (GRAD) {  // weights and their derivatives for the projection on the references
            const int ar = a + 4 * h;
            if (row && ar < NREF) {
              const double dgw = nz ? (dS[h] - gw * dnorm) * inv : 0.0;
              const double dzeta = -P.ga * gam * scale * zeta * qref[h] * qinv * qinv;
              const int o = i * NREF + ar;
              WT(WT_Q)[o] = wq[h];
              WT(WT_0)[o] = w0[h];
              WT(WT_ZGD)[o] = (T)(zeta * dgw);
              WT(WT_Z0GD)[o] = (T)(z0[h] * dgw);
              WT(WT_DZG)[o] = (T)(dzeta * gw);
            }
          }
        }
        if (!GRAD && A.alpha_out) {  // properties mode: alpha_i = sum_a zeta gw alpha_a(0)
          T al = wq[0] * tab.alpha0[za[0]] + (a + 4 < NREF ? wq[1] * tab.alpha0[za[1]] : T(0));
          al += __shfl_xor_sync(0xffffffffu, al, 2);
          al += __shfl_xor_sync(0xffffffffu, al, 1);
          if (a == 0 && row) A.alpha_out[(size_t)b * A.nat + idx[i]] = al;
        }
        // weighted polarizability vectors, both flavours: lane a takes w = a + 4 k
        const T* const al = tab.alpha_w + (size_t)z * NREF * NFREQ;
        T sq[6], s0[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) sq[k] = s0[k] = T(0);
#pragma unroll
        for (int ar = 0; ar < NREF; ++ar) {
          const T vq = __shfl_sync(0xffffffffu, wq[ar >> 2], (lane & 28) + (ar & 3));
          const T v0 = __shfl_sync(0xffffffffu, w0[ar >> 2], (lane & 28) + (ar & 3));
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            const int w = a + 4 * k;
            if (w < NFREQ) {
              const T av = al[ar * NFREQ + w];
              sq[k] += vq * av;
              s0[k] += v0 * av;
            }
          }
        }
        if (row) {
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            const int w = a + 4 * k;
            if (w < NFREQ) {
              Aq[w * AS + i] = sq[k];
              A0[w * AS + i] = s0[k];
            }
          }
        }
      }
    } else
#endif
    D4_ROWS8(i, a) {
      const T cn_row = row_sum2_8(pu, pu, i, a, n);  // every lane of the row holds the sum
      if (a == 0 && i < n) {
        if constexpr (GRAD || D4S) ATOM(AT_CN)[i] = cn_row;
        if (!GRAD && A.cn_out) A.cn_out[(size_t)b * A.nat + idx[i]] = cn_row;
      }
      const bool on = i < n && a < NREF;
      const int z = i < n ? zs[i] : 0;
      // branch-free body (one basic block): the four exponentials below are independent
      // chains that the scheduler interleaves; table entries are fetched up front
      const int za = z * NREF + (a < NREF ? a : 0);
      const int rc = on ? tab.refc[za] : 0;
      const double rcn = tab.refcn[za], qref = tab.refq[za], z0 = tab.zeta0[za];
      const double gam = tab.gamgc[z], zeff = tab.zeff[z];
      double qat = 0.0;
      if (i < n) {
        if constexpr (GRAD || D4S) qat = (double)ATOM(AT_Q)[i];
        else qat = (double)A.q[(size_t)b * A.nat + idx[i]];
      }
      const double d = (double)cn_row - rcn;
      const double arg = (rc > 0 && !D4S) ? P.wf * d * d : 1e300;
      double shift = arg;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) shift = fmin(shift, __shfl_xor_sync(0xffffffffu, shift, o));
      // sum_k exp(-(k arg - shift)) = exp(shift - arg) (1 + x + x^2 + ...), x = exp(-arg): two
      // exponentials per lane instead of one per Gaussian copy (refc is 1 or 3 in the D4 data)
      double S = 0.0, dS = 0.0;
      if constexpr (!D4S) {
        const double t1 = exp(shift - arg), x = exp(-arg);
        double pw = 1.0, acc = 0.0, dacc = 0.0;
#pragma unroll
        for (int k = 1; k <= 3; ++k) {
          acc += k <= rc ? pw : 0.0;
          dacc += k <= rc ? (double)k * pw : 0.0;
          pw *= x;
        }
        for (int k = 4; k <= rc; ++k) {
          acc += pw;
          dacc += (double)k * pw;
          pw *= x;
        }
        S = rc > 0 ? t1 * acc : 0.0;
        dS = rc > 0 ? -2.0 * P.wf * d * t1 * dacc : 0.0;
      }
      double norm = S, dnorm = dS;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        norm += __shfl_xor_sync(0xffffffffu, norm, o);
        dnorm += __shfl_xor_sync(0xffffffffu, dnorm, o);
      }
      // charge scaling zeta (model/base.py:326-335); qmod <= 0 is the scale = 0 limit
      const double qmod = qat + zeff;
      const bool qpos = qmod > 0.0;
      const double qe = qpos ? qmod - (double)d4_eps<T>() : 1.0;
      const double qinv = 1.0 / qe;
      const double scale = qpos ? exp(gam * (1.0 - qref * qinv)) : 0.0;
      const double zeta = rc > 0 ? exp(P.ga * (1.0 - scale)) : 0.0;
      const double dzeta = GRAD ? -P.ga * gam * scale * zeta * qref * qinv * qinv : 0.0;
      // normalised Gaussian weight; norm >= 1 where defined (the closest reference contributes exp(0))
      const bool nz = norm > 0.0;
      const double inv = d4_rcp(nz ? norm : 1.0);
      const double gw = nz ? S * inv : 0.0;
      const double dgw = nz ? (dS - gw * dnorm) * inv : 0.0;
      T wq_reg = T(0), w0_reg = T(0);  // zeta gw and zeta(q=0) gw of this lane's reference
      if (on) {
        const int o = i * NREF + a;
        wq_reg = (T)(zeta * gw);
        w0_reg = (T)(z0 * gw);
        if (D4S) {  // weights are pair dependent: keep the charge scaling only
          WT(WT_Q)[o] = (T)zeta;
          WT(WT_0)[o] = (T)z0;
          if (GRAD) WT(WT_ZGD)[o] = (T)dzeta;
        } else if (GRAD) {
          WT(WT_Q)[o] = wq_reg;
          WT(WT_0)[o] = w0_reg;
          WT(WT_ZGD)[o] = (T)(zeta * dgw);
          WT(WT_Z0GD)[o] = (T)(z0 * dgw);
          WT(WT_DZG)[o] = (T)(dzeta * gw);
        }
      }
      if constexpr (!GRAD && !D4S) {
        if (A.alpha_out) {  // properties mode: alpha_i = sum_a zeta gw alpha_a(0)
          T al = on ? wq_reg * tab.alpha0[za] : T(0);
          al += __shfl_xor_sync(0xffffffffu, al, 4);
          al += __shfl_xor_sync(0xffffffffu, al, 2);
          al += __shfl_xor_sync(0xffffffffu, al, 1);
          if (a == 0 && i < n) A.alpha_out[(size_t)b * A.nat + idx[i]] = al;
        }
      }
      if constexpr (!D4S) {
        // weighted polarizability vectors: lane a takes the frequencies a, a + 8, a + 16 and
        // collects the seven weights of its atom from the neighbouring lanes: the charge-scaled
        // flavour (two-body term) and the q = 0 flavour (ATM term) together.
        const T* const al = tab.alpha_w + (size_t)z * NREF * NFREQ;
        T sq[3] = {T(0), T(0), T(0)}, s0[3] = {T(0), T(0), T(0)};
#pragma unroll
        for (int ar = 0; ar < NREF; ++ar) {
          const T vq = __shfl_sync(0xffffffffu, wq_reg, (lane & 24) + ar);
          const T v0 = __shfl_sync(0xffffffffu, w0_reg, (lane & 24) + ar);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int w = a + 8 * k;
            if (w < NFREQ) {
              const T av = al[ar * NFREQ + w];
              sq[k] += vq * av;
              s0[k] += v0 * av;
            }
          }
        }
        if (i < n) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int w = a + 8 * k;
            if (w < NFREQ) {
              Aq[w * AS + i] = sq[k];
              A0[w * AS + i] = s0[k];
            }
          }
        }
      }
    }
    __syncthreads();
    PHASE(3);

    bool open = false;
    int nel = 0;           // D4S: distinct elements of the structure
    bool use_tab = false;  // D4S: weights come from the per-(atom, element) table
    T* const Rs = reinterpret_cast<T*>(smem + L::rs);  // D4S gradient: staged reference-C6 blocks
    if constexpr (D4S) {
      for (int i = tid; i < n; i += NT) atomicOr(&misc[8 + (zs[i] >> 5)], (int)(1u << (zs[i] & 31)));
      __syncthreads();
      nel = __popc(misc[8]) + __popc(misc[9]) + __popc(misc[10]) + __popc(misc[11]);
      use_tab = nel <= D4S_ECAP;
      if (use_tab) {
        for (int i = tid; i < n; i += NT) {
          const int z = zs[i], w = z >> 5;
          int e = __popc((unsigned)misc[8 + w] & ((1u << (z & 31)) - 1u));
          for (int ww = 0; ww < w; ++ww) e += __popc(misc[8 + ww]);
          ei[i] = e;
          misc[12 + e] = z;  // every atom of the element writes the same value
        }
        __syncthreads();
        if constexpr (GRAD) {
          if (nel <= D4S_RCAP) {
            for (int t = tid; t < nel * nel * (NREF * NREF); t += NT) {
              const int blk = t / (NREF * NREF), ab = t - blk * (NREF * NREF);
              const int e1 = blk / nel, e2 = blk - e1 * nel;
              Rs[t] = tab.rc6[((size_t)misc[12 + e1] * NELEM + misc[12 + e2]) * (NREF * NREF) + ab];
            }
          }
        }
        for (int t = tid; t < n * nel; t += NT) {
          const int i = t / nel, e = t - i * nel;
          const int zi = zs[i];
          T g[NREF], dg[NREF];
          d4s_weights<T, GRAD>(tab.refcn, tab.refc, zi, (double)ATOM(AT_CN)[i],
                               tab.wfpair[zi * NELEM + misc[12 + e]], g, dg);
#pragma unroll
          for (int a = 0; a < NREF; ++a) {
            wtab[(size_t)t * D4S_WSTR + a] = g[a];
            if (GRAD) wtab[(size_t)t * D4S_WSTR + NREF + a] = dg[a];
          }
        }
        __syncthreads();
      }
    }
    if constexpr (D4S) {
      // ---- D4S: pair-dependent Gaussian weights (model/d4s.py:109-290) ------
      // C6_ij = sum_ab rc6[Zi,Zj,a,b] (zeta_ia gw_ia|Zj) (zeta_jb gw_jb|Zi) with the weights
      // of each atom evaluated for the partner's element, per pair, in float64.
      // One fused pass: two-body energy (energy kernel) + ATM stash.
      for (int p = tid; p < np; p += NT) {
        int i, j;
        pair_lookup(tab.pij, p, i, j);
        const T r2 = pa[p];
        const int zi = zs[i], zj = zs[j];
        T gi[NREF], gj[NREF], dgu[NREF];
        if (use_tab) {
          const T* wj = wtab + (size_t)(j * nel + ei[i]) * D4S_WSTR;  // partner j as seen by i
          const T* wi = wtab + (size_t)(i * nel + ei[j]) * D4S_WSTR;
#pragma unroll
          for (int a = 0; a < NREF; ++a) {
            gj[a] = wj[a];
            gi[a] = wi[a];
          }
        } else {
          d4s_weights<T, false>(tab.refcn, tab.refc, zj, (double)ATOM(AT_CN)[j], tab.wfpair[zj * NELEM + zi], gj, dgu);
          d4s_weights<T, false>(tab.refcn, tab.refc, zi, (double)ATOM(AT_CN)[i], tab.wfpair[zi * NELEM + zj], gi, dgu);
        }
        T v[NREF], v0[NREF];
#pragma unroll
        for (int bq = 0; bq < NREF; ++bq) {
          v[bq] = WT(WT_Q)[j * NREF + bq] * gj[bq];
          v0[bq] = WT(WT_0)[j * NREF + bq] * gj[bq];
        }
        const T* R = (GRAD && use_tab && nel <= D4S_RCAP) ? Rs + (ei[i] * nel + ei[j]) * (NREF * NREF)
                                                          : tab.rc6 + ((size_t)zi * NELEM + zj) * (NREF * NREF);
        T c6q = T(0), c60 = T(0);
#pragma unroll
        for (int a = 0; a < NREF; ++a) {
          T t = T(0), t0 = T(0);
#pragma unroll
          for (int bq = 0; bq < NREF; ++bq) {
            const T rab = R[a * NREF + bq];
            t += rab * v[bq];
            t0 += rab * v0[bq];
          }
          const T g = gi[a];
          c6q += WT(WT_Q)[i * NREF + a] * g * t;
          c60 += WT(WT_0)[i * NREF + a] * g * t0;
        }
        const T ss = ATOM(AT_SQ)[i] * ATOM(AT_SQ)[j];
        const T R0 = P.a1 * ss + P.a2;
        if (!GRAD) {
          T e = T(0);
          if (r2 <= P.disp2_sq) {
            const T qq = ss * ss;
            const T r4 = r2 * r2, r6 = r4 * r2, r8 = r4 * r4;
            const T R2 = R0 * R0, R4 = R2 * R2, R6 = R4 * R2, R8 = R4 * R4;
            T F = P.s6 * d4_rcp(r6 + R6) + P.s8 * qq * d4_rcp(r8 + R8);
            if (P.s10k != T(0)) F += P.s10k * qq * qq * d4_rcp(r8 * r2 + R8 * R2);
            e = c6q * F;
          }
          out0[p] = e;
        }
        if (P.has_atm) {
          const T rinv = d4_rsqrt(r2);
          const bool inside = r2 <= P.disp3_sq;
          if (!inside) misc[2] = 1;
          pa[p] = inside ? r2 : -r2;
          pP[p] = d4_stash_p(P.fac9, c60, rinv);  // P' = P / r^2
          pu[p] = d4_zero_damp_arg(R0 * rinv, P.alp3, P.alp16 != 0);
        }
      }
      __syncthreads();
      PHASE(7);
      open = misc[2] != 0;
      if (!GRAD) {
        D4_ROWS8(i, sub) {
          const T e2 = row_sum2_8(out0, out0, i, sub, n);
          if (sub == 0 && i < n) ATOM(AT_E)[i] = T(-0.5) * e2;
        }
        __syncthreads();  // the triple loop reuses out0
      }
    }
    if constexpr (!D4S) {
    // ---- properties mode (disp.py:149-197): cn, C6_ij = A_i.A_j, alpha_i, then next structure
    if constexpr (!GRAD) {
      if (A.c6_out) {
        T* c6row = A.c6_out + (size_t)b * A.nat * A.nat;
        for (int t = tid; t < n * n; t += NT) {
          const int i = t / n, j = t - i * n;
          c6row[(size_t)idx[i] * A.nat + idx[j]] = dot23<T, AS>(Aq, i, j);
        }
        for (int i = tid; i < n; i += NT) A.energy[(size_t)b * A.nat + idx[i]] = T(0);
        continue;
      }
    }

    // ---- phase 4 (energy kernel): one pass over the pairs for the two-body energy
    // (twobody.py:134-201, rational damping) and the ATM pair stash (threebody.py:244-256,
    // 311-321).  FP64: both pair C6 flavours come from the tensor path (planes `pu`, `pP`).
    if constexpr (!GRAD) {
      T* const out2 = E2S ? Aq : out0 + 2 * CP;  // two-body pair energies (row sums in the final assembly)
      if constexpr (sizeof(T) == 8) {
        c6_tiles<AS>(Aq, pu, tab.pij, n, warp, lane, NW);
        if (P.has_atm) c6_tiles<AS>(A0, pP, tab.pij, n, warp, lane, NW);
        __syncthreads();
      }
#pragma unroll 2
      for (int p = tid; p < np; p += NT) {
        int i, j;
        pair_lookup(tab.pij, p, i, j);
        const T r2 = pa[p];
        const T ss = ATOM(AT_SQ)[i] * ATOM(AT_SQ)[j];
        const T R0 = P.a1 * ss + P.a2;
        T e = T(0);
        if (r2 <= P.disp2_sq) {
          T c6;
          if constexpr (sizeof(T) == 8) c6 = pu[p];
          else c6 = dot23<T, AS>(Aq, i, j);
          const T qq = ss * ss;  // = 3 r4r2_i r4r2_j
          const T r4 = r2 * r2, r6 = r4 * r2, r8 = r4 * r4;
          const T R2 = R0 * R0, R4 = R2 * R2, R6 = R4 * R2, R8 = R4 * R4;
          T F = P.s6 * d4_rcp(r6 + R6) + P.s8 * qq * d4_rcp(r8 + R8);
          if (P.s10k != T(0)) F += P.s10k * qq * qq * d4_rcp(r8 * r2 + R8 * R2);
          e = c6 * F;
        }
        out2[p] = e;
        if (P.has_atm) {
          T c60;
          if constexpr (sizeof(T) == 8) c60 = pP[p];
          else c60 = dot23<T, AS>(A0, i, j);
          const T rinv = d4_rsqrt(r2);
          const bool inside = r2 <= P.disp3_sq;
          if (!inside) misc[2] = 1;
          pa[p] = inside ? r2 : -r2;
          pP[p] = d4_stash_p(P.fac9, c60, rinv);  // P' = P / r^2
          pu[p] = d4_zero_damp_arg(R0 * rinv, P.alp3, P.alp16 != 0);
        }
      }
      __syncthreads();
      PHASE(5);
      open = misc[2] != 0;
    } else {
    // ---- phase 5 (gradient kernel): ATM pair stash --------------------------
    if constexpr (sizeof(T) == 8) {  // C6(q = 0) of all pairs on the tensor path -> plane `pP`
      if (P.has_atm) {
        c6_tiles<AS>(A0, pP, tab.pij, n, warp, lane, NW);
        __syncthreads();
      }
    }
    if (P.has_atm) {
#pragma unroll 2
      for (int p = tid; p < np; p += NT) {
        int i, j;
        pair_lookup(tab.pij, p, i, j);
        const T r2 = pa[p];
        const T rinv = d4_rsqrt(r2);
        T c6;
        if constexpr (sizeof(T) == 8) c6 = pP[p];
        else c6 = dot23<T, AS>(A0, i, j);
        const T R0 = P.a1 * ATOM(AT_SQ)[i] * ATOM(AT_SQ)[j] + P.a2;
        const bool inside = r2 <= P.disp3_sq;
        if (!inside) misc[2] = 1;
        pa[p] = inside ? r2 : -r2;
        pP[p] = d4_stash_p(P.fac9, c6, rinv);  // P' = P / r^2
        pu[p] = d4_zero_damp_arg(R0 * rinv, P.alp3, P.alp16 != 0);
      }
      __syncthreads();
      PHASE(7);
      open = misc[2] != 0;
    }
    }  // gradient kernel
    }  // !D4S
    if (P.has_atm) {
      // ---- phase 6: triple loop ---------------------------------------------
      if constexpr (GRAD) {
        // Gradient: one thread per owner pair (j,k), all third atoms i.  Every
        // triple is visited from each of its three pairs, so the per-pair sums
        // Gamma (dL/dC60 numerator) and D (dL/d r^2) need no communication.
        // Whole warps walk the pair list (lanes past the end repeat the last pair and skip the
        // stores) so that the sweep can use warp-uniform loop bounds.
        const T ifac9 = d4_rcp(P.fac9);
#ifndef D4_NO_TILED_SWEEP
        const bool tiled = !open && A.gin == nullptr && n >= 4;
#else
        const bool tiled = false;
#endif
        if (tiled)
          grad_sweep_tiled<T, D4S, CAP, NT>(pa, pP, pu, out0, out1, A.energy != nullptr, tab.pij, n, tid,
                                            T(0.5) * P.alp3, T(0.5) * P.alp3 - T(2.5), ifac9);
        else
        for (int p0 = warp * 32; p0 < np; p0 += NT) {
          const bool valid = p0 + lane < np;
          const int p = valid ? p0 + lane : np - 1;
          int j, k;
          pair_lookup(tab.pij, p, j, k);
          const T bs = pa[p];
          const T bb = fabs(bs);
          const T cjk = bs > T(0) ? T(1) : T(0);
          const T Pjk = pP[p], ujk = pu[p];
          const T inv_b = d4_rcp(bb), b2 = bb * bb, twob = bb + bb;
          const T gj = ATOM(AT_G)[j], gk = ATOM(AT_G)[k];
          const T gjk2 = T(2) * (gj + gk);
          const T kA = T(0.5) * P.alp3, kB = kA - T(2.5);  // alp / 6, alp / 6 - 2.5
          const bool pz = Pjk == T(0);  // no ATM contribution through this pair (C6(q=0) = 0)
          const T iP = pz ? T(1) : d4_rcp(Pjk);
          const T sPu = T(6) * ujk * iP, kAi = kA * iP;
          const int tj = j * (j - 1) / 2, tk = k * (k - 1) / 2;
          T accG = T(0), accC = T(0), accS = T(0), accH = T(0), accL = T(0);
          // branch-free sweep over the third atom: for i == j or i == k the visit runs on
          // the owner's own entry with a zero pair factor (contributes exactly 0), so the
          // body is straight-line code and four visits can be in flight per thread
#define D4_SWEEP(OPENV, UNITV, LO, HI)                                                          \
  _Pragma("unroll 4") for (int i = (LO); i < (HI); ++i) {                                       \
    const int ti = i * (i - 1) / 2;                                                             \
    const bool ok = (i != j) & (i != k);                                                        \
    const int pij = ok ? (i > j ? ti + j : tj + i) : p;                                         \
    const int pik = ok ? (i > k ? ti + k : tk + i) : p;                                         \
    grad_visit<T, OPENV, UNITV>(pa[pij], ok ? pP[pij] : T(0), pu[pij], pa[pik], pP[pik],        \
                                pu[pik], bb, b2, twob, cjk, iP, sPu, kAi, kB,                   \
                                UNITV ? T(0) : T(2) * ATOM(AT_G)[i], gjk2, gj, gk, accG, accC,  \
                                accS, accH, accL);                                              \
  }
          // closed structure, unit upstream weights: the two stash indices of a visit are affine
          // in i inside each of the ranges i < k, k < i < j, j < i -- no selects, no index
          // arithmetic beyond one running triangle offset
#define D4_VISIT_U(PIJ, PIK)                                                                     \
  grad_visit<T, false, true>(pa[PIJ], pP[PIJ], pu[PIJ], pa[PIK], pP[PIK], pu[PIK], bb, b2, twob,  \
                             cjk, iP, sPu, kAi, kB, T(0), gjk2, gj, gk, accG, accC, accS, accH, accL);
          if (open) {
            D4_SWEEP(true, false, 0, n)
          } else if (A.gin == nullptr) {
            const int jlo = __reduce_min_sync(0xffffffffu, j), jhi = __reduce_max_sync(0xffffffffu, j);
            if (jlo == jhi) {  // the warp's pairs share j (rows of 32 or more pairs: always)
              const int klo = __reduce_min_sync(0xffffffffu, k), khi = __reduce_max_sync(0xffffffffu, k);
#pragma unroll 4
              for (int i = 0; i < klo; ++i) {  // i < k < j
                D4_VISIT_U(tj + i, tk + i)
              }
              D4_SWEEP(false, true, klo, khi + 1)  // i crosses the k of some lanes
              int ti = (khi + 1) * khi / 2;
#pragma unroll 4
              for (int i = khi + 1; i < j; ++i) {  // k < i < j
                D4_VISIT_U(tj + i, ti + k)
                ti += i;
              }
              ti = (j + 1) * j / 2;
#pragma unroll 4
              for (int i = j + 1; i < n; ++i) {  // k < j < i
                D4_VISIT_U(ti + j, ti + k)
                ti += i;
              }
            } else {
              D4_SWEEP(false, true, 0, n)
            }
          } else {
            D4_SWEEP(false, false, 0, n)
          }
#undef D4_VISIT_U
#undef D4_SWEEP
          if (!open) {
            if (A.gin == nullptr) {  // unit upstream weights: W = 6, G is also the energy share
              accH = accG;
              accG *= T(6);
              accC *= T(6);
              accS *= T(6);
            }
            accH += accH;
            accL = accH;
          }
          if (pz) accG = accC = accS = accH = accL = T(0);
          const T accD = fma(accC, inv_b, T(0.375) * accS);
          if constexpr (!D4S) {
            // dL/dC6(q=0)_jk = Gamma / (2 C6): C6(q=0) = (P'_jk r^5 / fac9)^2 >= 0 is recovered from
            // the stash instead of a second 23-term dot product in the coefficient pass
            const T w = (Pjk * ifac9) * b2;
            accG = pz ? T(0) : accG * (T(0.5) * d4_rcp((w * w) * bb));
          }
          if (valid) {
            out0[p] = accG;
            out1[p] = accD;
            if (A.energy) {
              out0[2 * CP + p] = accH;
              out0[3 * CP + p] = accL;
            }
          }
        }
        __syncthreads();  // all reads of the stash done -> planes become outputs
        for (int p = tid; p < np; p += NT) {  // same thread wrote out0/out1[p]
          pP[p] = out0[p];
          pu[p] = out1[p];
        }
        __syncthreads();
        PHASE(8);
      } else {
        // Energy: every unordered triple i > j > k is evaluated exactly once.
        // A warp takes 32 consecutive "bottom" pairs (j,k) (one per lane, stash
        // entry in registers) and sweeps the top atom i in blocks of eight; the
        // shares of atoms j and k accumulate in lane registers, the share of
        // atom i is reduced over the lanes with a transposed shuffle reduction.
        T* const Tw = Aq + (E2S ? CP : 0) + warp * CAP;  // A vectors are dead: per-warp E_i partials
        for (int i = lane; i < n; i += 32) Tw[i] = T(0);
        __syncwarp();
        const int nchunks = (np + 31) >> 5;
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
        // static assignment of the chunks (sorted by decreasing sweep length): the
        // order of every floating-point sum is fixed, so results are bitwise reproducible
        for (int round = 0; round * NW < nchunks; ++round) {
          // boustrophedon: the sweep length decreases with the chunk index, so alternate
          // the direction in which the chunks of a round are dealt to the warps
          const int chunk = round * NW + ((round & 1) ? NW - 1 - warp : warp);
          if (chunk >= nchunks) continue;
          const int p = chunk * 32 + lane;
          const bool valid = p < np;
          int j = 1 << 20, k = 0;
          if (valid) pair_lookup(tab.pij, p, j, k);
          const int jmin = __shfl_sync(0xffffffffu, j, 0);
          // last row that is only partially active; an invalid lane keeps the
          // whole sweep on the predicated path
          const int jmax = __reduce_max_sync(0xffffffffu, j);
          const T bs = valid ? pa[p] : T(1);
          const T Pjk = valid ? pP[p] : T(0), ujk = valid ? pu[p] : T(0);
          const T bb = fabs(bs), bb2 = bb * bb;
          const T cjk = bs > T(0) ? T(1) : T(0);
          const T* colj = pa + (valid ? j : 0);
          const T* colk = pa + k;
          T accJ = T(0), accK = T(0);
          T v[8];
          for (int i0 = jmin + 1; i0 < n; i0 += 8) {
            const bool all = i0 > jmax && i0 + 8 <= n;  // warp-uniform
            if (open)
              triple_block8<T, CP, true, true>(colj, colk, i0, j, n, bb, bb2, cjk, Pjk, ujk, accJ, accK, v);
            else if (all)
              triple_block8<T, CP, false, false>(colj, colk, i0, j, n, bb, bb2, cjk, Pjk, ujk, accJ, accK, v);
            else
              triple_block8<T, CP, true, false>(colj, colk, i0, j, n, bb, bb2, cjk, Pjk, ujk, accJ, accK, v);
            const T y = reduce8x32(v, b4, b3, b2);
            const int iw = i0 + (lane >> 2);
            if ((lane & 3) == 0 && iw < n) Tw[iw] += y;
          }
          // shares of the bottom pairs -> per-warp partials: row by row of the chunk (the
          // lanes of one row j have distinct k), the j share as a warp sum
          {
            const T shareK = open ? accK : accJ;
            const int jtop = __reduce_max_sync(0xffffffffu, valid ? j : 0);
            for (int jj = jmin; jj <= jtop; ++jj) {
              const bool mine = valid && j == jj;
              const T sj = warp_sum(mine ? accJ : T(0));
              if (mine) Tw[k] += shareK;
              if (lane == 0) Tw[jj] += sj;
              __syncwarp();
            }
          }
        }
        __syncthreads();
        PHASE(8);
      }
    } else if (GRAD) {
      for (int p = tid; p < np; p += NT) {
        pP[p] = T(0);
        pu[p] = T(0);
      }
      __syncthreads();
    }

    if constexpr (!GRAD) {
      // ---- final assembly: E_i = E2_i + scale * (ATM shares) ------------------
      const T scale = open ? T(1) : T(2);  // closed triples: every atom has multiplicity 2
      const T* const e2p = E2S ? Aq : out0 + 2 * CP;
      const T* const Tall = Aq + (E2S ? CP : 0);
      D4_ROWS8(i, sub) {
        T e = T(0);
        if constexpr (!D4S) e = T(-0.5) * row_sum2_8(e2p, e2p, i, sub, n);
        if (sub == 0 && i < n) {
          if constexpr (D4S) e = ATOM(AT_E)[i];
          if (P.has_atm) {
            T s3 = T(0);
#pragma unroll
            for (int w = 0; w < NW; ++w) s3 += Tall[w * CAP + i];
            e += scale * s3;
          }
          A.energy[(size_t)b * A.nat + idx[i]] = e;
        }
      }
      __syncthreads();
      PHASE(9);
    } else {
      // =================== gradient back-propagation =========================
      if constexpr (D4S) {
        // D4S: weights (and their cn-derivatives) are re-evaluated per pair; the four
        // chain-rule contributions of a pair go to per-pair planes that are summed
        // per atom afterwards.
        //   pa <- d L/d cn (share of the higher-index atom)   pP <- (lower-index atom)
        //   out0/out1 <- d L/d q shares                       pu <- radial force coefficient
        for (int p = tid; p < np; p += NT) {
          int i, j;
          pair_lookup(tab.pij, p, i, j);
          const T r2 = fabs(pa[p]);  // stash: signed squared distance
          const int zi = zs[i], zj = zs[j];
          const T* R = (use_tab && nel <= D4S_RCAP) ? Rs + (ei[i] * nel + ei[j]) * (NREF * NREF)
                                                    : tab.rc6 + ((size_t)zi * NELEM + zj) * (NREF * NREF);
          // weights (gw[7], d gw/d cn[7]) of i as seen by j's element and vice versa: from the per-(atom,
          // element) table, or evaluated here when the structure has more distinct elements than it holds
          T wloc[2][2 * NREF];
          const T* wi = wloc[0];
          const T* wj = wloc[1];
          if (use_tab) {
            wi = wtab + (size_t)(i * nel + ei[j]) * D4S_WSTR;
            wj = wtab + (size_t)(j * nel + ei[i]) * D4S_WSTR;
          } else {
            T g[NREF], dg[NREF];
            d4s_weights<T, true>(tab.refcn, tab.refc, zi, (double)ATOM(AT_CN)[i], tab.wfpair[zi * NELEM + zj], g, dg);
#pragma unroll
            for (int a = 0; a < NREF; ++a) wloc[0][a] = g[a], wloc[0][NREF + a] = dg[a];
            d4s_weights<T, true>(tab.refcn, tab.refc, zj, (double)ATOM(AT_CN)[j], tab.wfpair[zj * NELEM + zi], g, dg);
#pragma unroll
            for (int a = 0; a < NREF; ++a) wloc[1][a] = g[a], wloc[1][NREF + a] = dg[a];
          }
          // One SIDE at a time (x = i with partner j, then x = j with partner i, R transposed): t = R_xy v_y
          // with v_y = zeta_y o gw_y|Zx, then everything that differentiates x's weights.  Two passes over
          // the 7 x 7 block with ~30 live values each instead of one pass with ~60 (which spilled).
          const T* Rji = (use_tab && nel <= D4S_RCAP) ? Rs + (ei[j] * nel + ei[i]) * (NREF * NREF)
                                                      : tab.rc6 + ((size_t)zj * NELEM + zi) * (NREF * NREF);
          T c6q = T(0), c60 = T(0), dq_cni = T(0), d0_cni = T(0), dq_qi = T(0);
          T dq_cnj = T(0), d0_cnj = T(0), dq_qj = T(0);
#pragma unroll
          for (int side = 0; side < 2; ++side) {
            const int x = side ? j : i, y = side ? i : j;
            const T* Rxy = side ? Rji : R;
            const T* wx = side ? wj : wi;
            const T* wy = side ? wi : wj;
            T vq[NREF], v0[NREF];
#pragma unroll
            for (int bq = 0; bq < NREF; ++bq) {
              const T g = wy[bq];
              vq[bq] = WT(WT_Q)[y * NREF + bq] * g;
              v0[bq] = WT(WT_0)[y * NREF + bq] * g;
            }
            T cq = T(0), c0 = T(0), dcq = T(0), dc0 = T(0), dqq = T(0);
#pragma unroll
            for (int a = 0; a < NREF; ++a) {
              T t = T(0), t0 = T(0);
#pragma unroll
              for (int bq = 0; bq < NREF; ++bq) {
                const T rab = Rxy[a * NREF + bq];
                t += rab * vq[bq];
                t0 += rab * v0[bq];
              }
              const T gx = wx[a], dgx = wx[NREF + a];
              const T zq = WT(WT_Q)[x * NREF + a], z0 = WT(WT_0)[x * NREF + a];
              cq += zq * gx * t;
              c0 += z0 * gx * t0;
              dcq += zq * dgx * t;
              dc0 += z0 * dgx * t0;
              dqq += WT(WT_ZGD)[x * NREF + a] * gx * t;
            }
            if (side == 0) {
              c6q = cq, c60 = c0, dq_cni = dcq, d0_cni = dc0, dq_qi = dqq;
            } else {
              dq_cnj = dcq, d0_cnj = dc0, dq_qj = dqq;
            }
          }
          const T G2 = T(-0.5) * (ATOM(AT_G)[i] + ATOM(AT_G)[j]);
          T coefq = T(0), fc = T(2) * pu[p], e2 = T(0);
          if (r2 <= P.disp2_sq) {
            const T ss = ATOM(AT_SQ)[i] * ATOM(AT_SQ)[j];
            const T R0 = P.a1 * ss + P.a2;
            const T qq = ss * ss;
            const T r4 = r2 * r2, r6 = r4 * r2, r8 = r4 * r4;
            const T R2 = R0 * R0, R4 = R2 * R2, R6 = R4 * R2, R8 = R4 * R4;
            const T t6 = d4_rcp(r6 + R6), t8 = d4_rcp(r8 + R8);
            T F = P.s6 * t6 + P.s8 * qq * t8;
            T dF = -(T(6) * P.s6 * r4 * t6 * t6 + T(8) * P.s8 * qq * r6 * t8 * t8);
            if (P.s10k != T(0)) {
              const T t10 = d4_rcp(r8 * r2 + R8 * R2);
              F += P.s10k * qq * qq * t10;
              dF -= T(10) * P.s10k * qq * qq * r8 * t10 * t10;
            }
            coefq = G2 * F;
            fc += G2 * c6q * dF;
            e2 = c6q * F;
          }
          if (A.energy) out0[4 * CP + p] = e2;
          const T gam = c60 != T(0) ? pP[p] / (T(2) * c60) : T(0);
          pa[p] = coefq * dq_cni + gam * d0_cni;
          pP[p] = coefq * dq_cnj + gam * d0_cnj;
          out0[p] = coefq * dq_qi;
          out1[p] = coefq * dq_qj;
          pu[p] = fc;
        }
        __syncthreads();
        PHASE(10);
        D4_ROWS8(i, sub) {
          const T dc = row_sum2_8(pa, pP, i, sub, n);
          const T dq = row_sum2_8(out0, out1, i, sub, n);
          if (sub == 0 && i < n) {
            ATOM(AT_DCN)[i] = dc;
            ATOM(AT_DQ)[i] = dq;
          }
        }
        __syncthreads();
        PHASE(12);
      } else {
      // phase 7: per-pair coefficients
      //   pa <- G2 F          (dL/dC6q)
      //   pP =  Gamma/(2 C60) (dL/dC60, from the sweep)
      //   pu <- 2 D + G2 C6q F'/r   (radial force coefficient, CN chain added later)
      if constexpr (sizeof(T) == 8) {  // C6(q) of all pairs on the tensor path -> scratch plane out1 (free again)
        c6_tiles<AS>(Aq, out1, tab.pij, n, warp, lane, NW);
        __syncthreads();
      }
      for (int p = tid; p < np; p += NT) {
        int i, j;
        pair_lookup(tab.pij, p, i, j);
        const T r2 = fabs(pa[p]);  // stash: signed squared distance
        T c6q;
        if constexpr (sizeof(T) == 8) c6q = out1[p];
        else c6q = dot23<T, AS>(Aq, i, j);
        const T G2 = T(-0.5) * (ATOM(AT_G)[i] + ATOM(AT_G)[j]);
        T coefq = T(0), fc = T(2) * pu[p], e2 = T(0);
        if (r2 <= P.disp2_sq) {
          const T ss = ATOM(AT_SQ)[i] * ATOM(AT_SQ)[j];
          const T R0 = P.a1 * ss + P.a2;
          const T qq = ss * ss;
          const T r4 = r2 * r2, r6 = r4 * r2, r8 = r4 * r4;
          const T R2 = R0 * R0, R4 = R2 * R2, R6 = R4 * R2, R8 = R4 * R4;
          const T t6 = d4_rcp(r6 + R6), t8 = d4_rcp(r8 + R8);
          T F = P.s6 * t6 + P.s8 * qq * t8;
          // dF/dr / r
          T dF = -(T(6) * P.s6 * r4 * t6 * t6 + T(8) * P.s8 * qq * r6 * t8 * t8);
          if (P.s10k != T(0)) {
            const T t10 = d4_rcp(r8 * r2 + R8 * R2);
            F += P.s10k * qq * qq * t10;
            dF -= T(10) * P.s10k * qq * qq * r8 * t10 * t10;
          }
          coefq = G2 * F;
          fc += G2 * c6q * dF;
          e2 = c6q * F;
        }
        if (A.energy) out0[4 * CP + p] = e2;
        pa[p] = coefq;  // pP already holds Gamma / (2 C60) (scaled at the end of the sweep)
        pu[p] = fc;
      }
      __syncthreads();
      PHASE(10);
      // phase 8: B_i[w] = sum_j coef_ij A_j[w]: an (n x n).(n x 23) matrix product per flavour
      if constexpr (sizeof(T) == 8) {
        // FP64 tensor path (mma.m8n8k4): a warp owns an 8-atom x 8-frequency output tile
        // and walks the contraction index j in steps of four.  Operand fragments:
        //   A[r][c] = coef(i0 + r, j0 + c), B[c][r] = A_{j0 + c}[w0 + r], r = lane / 4, c = lane % 4
        const int nrb = (n + 7) >> 3;
        const int r = lane >> 2, c = lane & 3;
        // both flavours of a tile in the same warp: two independent accumulator chains share the index
        // arithmetic (the chain of dependent DMMAs is latency bound on its own)
        for (int tile = warp; tile < nrb * 3; tile += NW) {
          const int wb = tile % 3, rb = tile / 3;
          const int i = rb * 8 + r, ti = i * (i - 1) / 2;
          const int w = wb * 8 + r;
          const bool iok = i < n, wok = w < NFREQ;
          double q0 = 0.0, q1 = 0.0, z0 = 0.0, z1 = 0.0;
          for (int j0 = 0; j0 < n; j0 += 4) {
            const int j = j0 + c;
            const bool jok = j < n;
            double aq = 0.0, a0 = 0.0, bq = 0.0, b0 = 0.0;
            if (iok && jok && j != i) {
              const int pp = j < i ? ti + j : j * (j - 1) / 2 + i;
              aq = pa[pp];
              a0 = pP[pp];
            }
            if (wok && jok) {
              bq = Aq[w * AS + j];
              b0 = A0[w * AS + j];
            }
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(q0), "+d"(q1)
                         : "d"(aq), "d"(bq));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(z0), "+d"(z1)
                         : "d"(a0), "d"(b0));
          }
          // accumulator fragment: row r (atom), columns 2c, 2c+1 (frequency)
          const int wo = wb * 8 + 2 * c;
          if (iok) {
            if (wo < NFREQ) {
              Bq[wo * AS + i] = q0;
              B0[wo * AS + i] = z0;
            }
            if (wo + 1 < NFREQ) {
              Bq[(wo + 1) * AS + i] = q1;
              B0[(wo + 1) * AS + i] = z1;
            }
          }
        }
      } else {
        for (int t = tid; t < NFREQ * n; t += NT) {
          const int w = t / n, i = t - w * n;
          const int ti = i * (i - 1) / 2;
          T sq = T(0), s0 = T(0);
          for (int j = 0; j < i; ++j) {
            sq += pa[ti + j] * Aq[w * AS + j];
            s0 += pP[ti + j] * A0[w * AS + j];
          }
          for (int j = i + 1; j < n; ++j) {
            const int pj = j * (j - 1) / 2 + i;
            sq += pa[pj] * Aq[w * AS + j];
            s0 += pP[pj] * A0[w * AS + j];
          }
          Bq[w * AS + i] = sq;
          B0[w * AS + i] = s0;
        }
      }
      __syncthreads();
      PHASE(11);
      // phase 9: project on the references -> dL/dcn_i, dL/dq_i
      T* const tcn = pa;              // [7n] partial dL/dcn (pa is free again)
      T* const tq = pa + NREF * CAP;  // [7n] partial dL/dq
      for (int t = tid; t < NREF * n; t += NT) {
        const int i = t / NREF, a = t - i * NREF;
        const T* al = tab.alpha_w + ((size_t)zs[i] * NREF + a) * NFREQ;
        T pq = T(0), p0 = T(0);
#pragma unroll
        for (int w = 0; w < NFREQ; ++w) {
          const T av = al[w];
          pq += av * Bq[w * AS + i];
          p0 += av * B0[w * AS + i];
        }
        tcn[t] = WT(WT_ZGD)[t] * pq + WT(WT_Z0GD)[t] * p0;
        tq[t] = WT(WT_DZG)[t] * pq;
      }
      __syncthreads();
      for (int i = tid; i < n; i += NT) {
        T sc = T(0), sq = T(0);
#pragma unroll
        for (int a = 0; a < NREF; ++a) {
          sc += tcn[i * NREF + a];
          sq += tq[i * NREF + a];
        }
        ATOM(AT_DCN)[i] = sc;
        ATOM(AT_DQ)[i] = sq;
      }
      __syncthreads();
      PHASE(12);
      if (tid == 0) prefetch_next();  // the B vectors are dead: they receive the next structure's rows
      }  // D4 / D4S
      // phase 10: CN chain rule, d cn/d r = -den kcn/(r0 sqrt(pi)) exp(-x^2)
      for (int p = tid; p < np; p += NT) {
        int i, j;
        pair_lookup(tab.pij, p, i, j);
        const T dx = ATOM(AT_X)[i] - ATOM(AT_X)[j];
        const T dy = ATOM(AT_Y)[i] - ATOM(AT_Y)[j];
        const T dz = ATOM(AT_Z)[i] - ATOM(AT_Z)[j];
        const T r2 = dx * dx + dy * dy + dz * dz;
        if (r2 <= P.cn_sq) {
          const T rinv = d4_rsqrt(r2);
          const T r = r2 * rinv;
          const T r0inv = d4_rcp(ATOM(AT_RCOV)[i] + ATOM(AT_RCOV)[j]);
          const T xx = T(7.5) * (r * r0inv - T(1));
          if (fabs(xx) < T(8.7)) {  // exp(-x^2) < 1e-32 beyond
            const T dcn = -tab.den[zs[i] * NELEM + zs[j]] * T(7.5) * T(0.5641895835477563) *
                          r0inv * d4_exp(-xx * xx);
            pu[p] += (ATOM(AT_DCN)[i] + ATOM(AT_DCN)[j]) * dcn * rinv;
          }
        }
        if (A.energy) {
          // fused energy + gradient call: per-pair energy shares of the higher- / lower-index atom
          // from the L2 scratch (coalesced) into the planes `pa` / `pP`, which are free by now; the
          // force gather below sums them along the rows (no strided column reads from global memory)
          const T e2 = T(-0.5) * out0[4 * CP + p];
          pa[p] = P.has_atm ? fma(T(0.5), out0[2 * CP + p], e2) : e2;
          pP[p] = P.has_atm ? fma(T(0.5), out0[3 * CP + p], e2) : e2;
        }
      }
      __syncthreads();
      PHASE(13);
      // phase 11: gather forces, eight lanes per atom (four atoms per warp: all warps busy for a few
      // short passes instead of a long chain per atom), fixed summation order
      {
        const bool want_e = A.energy != nullptr;
        D4_ROWS8(i, sub) {
          T fx = T(0), fy = T(0), fz = T(0), es = T(0);
          if (i < n) {
            const T xi = ATOM(AT_X)[i], yi = ATOM(AT_Y)[i], zi = ATOM(AT_Z)[i];
            const int ti = i * (i - 1) / 2;
            for (int j = sub; j < n; j += 8) {
              if (j == i) continue;
              const int pp = j < i ? ti + j : j * (j - 1) / 2 + i;
              const T c = pu[pp];
              fx += c * (xi - ATOM(AT_X)[j]);
              fy += c * (yi - ATOM(AT_Y)[j]);
              fz += c * (zi - ATOM(AT_Z)[j]);
              if (want_e) es += j < i ? pa[pp] : pP[pp];
            }
          }
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) {
            fx += __shfl_xor_sync(0xffffffffu, fx, o);
            fy += __shfl_xor_sync(0xffffffffu, fy, o);
            fz += __shfl_xor_sync(0xffffffffu, fz, o);
            es += __shfl_xor_sync(0xffffffffu, es, o);
          }
          if (sub == 0 && i < n) {
            const size_t o = (size_t)b * A.nat + idx[i];
            if (want_e) A.energy[o] = es;
            if (A.grad) {
              A.grad[3 * o] = fx;
              A.grad[3 * o + 1] = fy;
              A.grad[3 * o + 2] = fz;
            }
            if (A.gradq) A.gradq[o] = ATOM(AT_DQ)[i];
          }
        }
      }
      __syncthreads();
      PHASE(14);
    }
  }
#undef PHASE
#undef ATOM
#undef WT
}

}  // namespace d4b200

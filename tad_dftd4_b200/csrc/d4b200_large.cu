// Tiled kernels for single large structures (hundreds to tens of thousands of
// atoms): the N^2 / N^3 tensors of the reference (README.md:355-357 of the reference:
// "the ATM term requires a 3D tensor (n_atoms, n_atoms, n_atoms)") never exist.
//
//   large_cn        CN of every atom (tile sweep over all columns, hard cutoff 30 Bohr)
//   large_weights   Gaussian weights x zeta  ->  per-atom weights W_q, W_0
//   large_avec      weighted polarizability vectors A_q[N][24], A_0[N][24] (C6_ij = A_i . A_j)
//   large_twobody   E_i = -1/2 sum_j C6q_ij F_ij for the rows of this rank (cutoff 60)
//   large_union     per group of 16 consecutive centre atoms: sorted list of all atoms
//                   within the ATM cutoff (40) of any centre + 16-bit membership mask
//   large_atm       centre-based ATM sum (reproduces the reference's two-distance mask,
//                   threebody.py:153-157, by construction): for every centre j and every
//                   unordered pair {i,k} of its neighbours, E_i += e/6 and E_k += e/6
//
// Atoms are expected in a spatially coherent order (the Python front end sorts them
// along a Morton curve); correctness does not depend on it.  Multi-GPU: ranks own
// disjoint ranges of rows (two-body) and centre groups (ATM) and all-reduce the
// per-atom energies; CN, weights and A vectors are recomputed on every rank (cheaper
// than a gather).
#include <cuda_runtime.h>
#include <math.h>

#include "d4b200_handle.cuh"
#include <cstdint>
#include <cstdlib>

#include "d4b200_small.cuh"  // math shims, d4_rcp, d4_zero_damp_arg, warp_sum

namespace d4b200 {

#ifndef D4_LARGE_PIPE_DEFAULT
#define D4_LARGE_PIPE_DEFAULT true  // large_atm_grad_pipe instead of large_atm_grad (D4B200_LARGE_PIPE=0/1 overrides)
#endif
#ifndef D4_LARGE_GROUP
#define D4_LARGE_GROUP 16
#endif
constexpr int GROUP = D4_LARGE_GROUP;   // centres per ATM group (<= 32: one mask bit each)
constexpr int TILE = 32;    // atoms per tile of the union list
constexpr int AVEC = 24;    // padded length of a polarizability vector

template <typename T>
struct LargeArgs {
  const int64_t* numbers;  // [nat]
  const T* pos;            // [nat,3]
  const T* q;              // [nat]
  T* energy;               // [nat] partial energies of this rank (accumulated)
  T* cn;                   // [nat]
  T* wq;                   // [nat,7]
  T* w0;                   // [nat,7]
  T* aq;                   // [nat,AVEC]
  T* a0;                   // [nat,AVEC]
  int* ulist;              // [ngroups, nat] union neighbour lists
  unsigned* umask;         // [ngroups, nat] membership masks (bit j: within cutoff of centre j)
  int* ucount;             // [ngroups]
  T* cstash;               // [gridDim.x][GROUP][3][ucap] per-CTA centre stash scratch
  int* queue;              // [1] dynamic group counter
  int* status;
  // gradient path
  const T* gin;            // [nat] upstream dL/dE (nullable = ones)
  T* zgd;                  // [nat,7] zeta * dgw/dcn
  T* z0gd;                 // [nat,7] zeta0 * dgw/dcn
  T* dzg;                  // [nat,7] dzeta/dq * gw
  T* daq_cn;               // [nat,AVEC] dA_q/dcn
  T* da0_cn;               // [nat,AVEC] dA_0/dcn
  T* daq_q;                // [nat,AVEC] dA_q/dq
  T* force;                // [nat,3] accumulated dL/dR
  T* dcn;                  // [nat] accumulated dL/dcn (stage 2 input: total over ranks)
  T* dq;                   // [nat] accumulated dL/dq
  int nat, ngroups, ucap;
  int row_begin, row_end;      // two-body rows of this rank
  int group_begin, group_end;  // ATM centre groups of this rank
  int nslice;                  // work item = (group, slice of its row tiles): nslice items per group
  Tables<T> tab;
  Par<T> par;
};

// --------------------------------------------------------------------- CN
template <typename T>
__global__ void __launch_bounds__(128) large_cn(LargeArgs<T> A, int ncolchunks) {
  __shared__ T sx[128], sy[128], sz[128], sr[128];
  __shared__ int szn[128];
  const int i = blockIdx.x * 128 + threadIdx.x;
  const bool act = i < A.nat;
  const int zi = act ? (int)A.numbers[i] : 0;
  const T xi = act ? A.pos[3 * i] : T(0), yi = act ? A.pos[3 * i + 1] : T(0), zi_ = act ? A.pos[3 * i + 2] : T(0);
  const T ri = act && zi > 0 && zi < NELEM ? A.tab.rcov[zi] : T(1);
  if (act && (zi < 0 || zi >= NELEM)) atomicOr(A.status, D4B200_STATUS_BAD_NUMBER);
  const int chunk = (A.nat + ncolchunks - 1) / ncolchunks;
  const int c0 = blockIdx.y * chunk, c1 = min(A.nat, c0 + chunk);
  T acc = T(0);
  for (int base = c0; base < c1; base += 128) {
    const int j = base + threadIdx.x;
    __syncthreads();
    if (j < c1) {
      const int zj = (int)A.numbers[j];
      sx[threadIdx.x] = A.pos[3 * j];
      sy[threadIdx.x] = A.pos[3 * j + 1];
      sz[threadIdx.x] = A.pos[3 * j + 2];
      szn[threadIdx.x] = zj > 0 && zj < NELEM ? zj : 0;
      sr[threadIdx.x] = zj > 0 && zj < NELEM ? A.tab.rcov[zj] : T(1);
    }
    __syncthreads();
    const int m = min(128, c1 - base);
    if (act && zi > 0 && zi < NELEM) {
      for (int t = 0; t < m; ++t) {
        const T dx = xi - sx[t], dy = yi - sy[t], dz = zi_ - sz[t];
        const T r2 = dx * dx + dy * dy + dz * dz;
        if (r2 <= A.par.cn_sq && base + t != i && szn[t] != 0) {
          const T r = d4_sqrt(r2);
          const T xx = T(7.5) * (r * d4_rcp(ri + sr[t]) - T(1));
          if (xx < d4_erfc_cut(T(0))) acc += A.tab.den[zi * NELEM + szn[t]] * T(0.5) * d4_erfc(xx);
        }
      }
    }
  }
  if (act && acc != T(0)) atomicAdd(&A.cn[i], acc);
}

// ---------------------------------------------------------------- weights
template <typename T, bool GRAD>
__global__ void __launch_bounds__(256) large_weights(LargeArgs<T> A) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  const int i = t >> 3, a = t & 7;
  const int zraw = i < A.nat ? (int)A.numbers[i] : 0;
  const bool on = i < A.nat && a < NREF && zraw > 0 && zraw < NELEM;
  const int z = on ? zraw : 0;
  const int rc = on ? A.tab.refc[z * NREF + a] : 0;
  const double d = on ? (double)A.cn[i] - A.tab.refcn[z * NREF + a] : 0.0;
  const double arg = rc > 0 ? A.par.wf * d * d : 1e300;
  double shift = arg;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) shift = fmin(shift, __shfl_xor_sync(0xffffffffu, shift, o));
  double S = 0.0, dS = 0.0;
  for (int k = 1; k <= rc; ++k) {
    const double e = exp(-((double)k * arg - shift));
    S += e;
    dS += -2.0 * (double)k * A.par.wf * d * e;
  }
  double norm = S, dnorm = dS;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    norm += __shfl_xor_sync(0xffffffffu, norm, o);
    dnorm += __shfl_xor_sync(0xffffffffu, dnorm, o);
  }
  if (i < A.nat && a < NREF) {
    double gw = 0.0, dgw = 0.0, zeta = 0.0, dzeta = 0.0;
    if (norm > 0.0) {
      gw = S / norm;
      dgw = (dS - gw * dnorm) / norm;
    }
    if (rc > 0) {
      const double qmod = (double)A.q[i] + A.tab.zeff[z];
      if (qmod > 0.0) {
        const double gam = A.tab.gamgc[z], qref = A.tab.refq[z * NREF + a];
        const double qe = qmod - (double)d4_eps<T>();
        const double scale = exp(gam * (1.0 - qref / qe));
        zeta = exp(A.par.ga * (1.0 - scale));
        dzeta = -A.par.ga * gam * scale * zeta * qref / (qe * qe);
      } else {
        zeta = exp(A.par.ga);
      }
    }
    const double z0 = on ? A.tab.zeta0[z * NREF + a] : 0.0;
    A.wq[i * NREF + a] = (T)(zeta * gw);
    A.w0[i * NREF + a] = (T)(z0 * gw);
    if (GRAD) {
      A.zgd[i * NREF + a] = (T)(zeta * dgw);
      A.z0gd[i * NREF + a] = (T)(z0 * dgw);
      A.dzg[i * NREF + a] = (T)(dzeta * gw);
    }
  }
}

template <typename T, bool GRAD>
__global__ void __launch_bounds__(256) large_avec(LargeArgs<T> A) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  const int i = t / AVEC, w = t - i * AVEC;
  if (i >= A.nat) return;
  T sq = T(0), s0 = T(0), dqc = T(0), d0c = T(0), dqq = T(0);
  const int zraw = (int)A.numbers[i];
  if (w < NFREQ && zraw > 0 && zraw < NELEM) {
    const T* al = A.tab.alpha_w + (size_t)zraw * NREF * NFREQ + w;
#pragma unroll
    for (int a = 0; a < NREF; ++a) {
      const T av = al[a * NFREQ];
      sq += A.wq[i * NREF + a] * av;
      s0 += A.w0[i * NREF + a] * av;
      if (GRAD) {
        dqc += A.zgd[i * NREF + a] * av;
        d0c += A.z0gd[i * NREF + a] * av;
        dqq += A.dzg[i * NREF + a] * av;
      }
    }
  }
  A.aq[(size_t)i * AVEC + w] = sq;
  A.a0[(size_t)i * AVEC + w] = s0;
  if (GRAD) {
    A.daq_cn[(size_t)i * AVEC + w] = dqc;
    A.da0_cn[(size_t)i * AVEC + w] = d0c;
    A.daq_q[(size_t)i * AVEC + w] = dqq;
  }
}

// ---------------------------------------------------------------- two-body
// One thread per row atom i (its A_q vector in registers), column tiles staged in
// shared memory; lanes of a warp read the same column -> broadcast loads.
template <typename T>
__global__ void __launch_bounds__(128) large_twobody(LargeArgs<T> A, int ncolchunks) {
  __shared__ T sA[64 * AVEC];
  __shared__ T sx[64], sy[64], sz[64], ss[64];
  __shared__ int sreal[64];
  const int i = A.row_begin + blockIdx.x * 128 + threadIdx.x;
  const bool act = i < A.row_end;
  const int zi = act ? (int)A.numbers[i] : 0;
  const bool real_i = act && zi > 0 && zi < NELEM;
  T ai[NFREQ];
#pragma unroll
  for (int w = 0; w < NFREQ; ++w) ai[w] = real_i ? A.aq[(size_t)i * AVEC + w] : T(0);
  const T xi = act ? A.pos[3 * i] : T(0), yi = act ? A.pos[3 * i + 1] : T(0), zi_ = act ? A.pos[3 * i + 2] : T(0);
  const T si = real_i ? A.tab.sqrt_r4r2[zi] : T(0);
  const Par<T>& P = A.par;
  const int chunk = (A.nat + ncolchunks - 1) / ncolchunks;
  const int c0 = blockIdx.y * chunk, c1 = min(A.nat, c0 + chunk);
  T acc = T(0);
  for (int base = c0; base < c1; base += 64) {
    __syncthreads();
    for (int t = threadIdx.x; t < 64 * AVEC; t += 128) {
      const int j = base + t / AVEC;
      sA[t] = j < c1 ? A.aq[(size_t)j * AVEC + (t % AVEC)] : T(0);
    }
    if (threadIdx.x < 64) {
      const int j = base + threadIdx.x;
      const int zj = j < c1 ? (int)A.numbers[j] : 0;
      const bool rj = zj > 0 && zj < NELEM;
      sx[threadIdx.x] = j < c1 ? A.pos[3 * j] : T(0);
      sy[threadIdx.x] = j < c1 ? A.pos[3 * j + 1] : T(0);
      sz[threadIdx.x] = j < c1 ? A.pos[3 * j + 2] : T(0);
      ss[threadIdx.x] = rj ? A.tab.sqrt_r4r2[zj] : T(0);
      sreal[threadIdx.x] = rj;
    }
    __syncthreads();
    if (!real_i) continue;
    const int m = min(64, c1 - base);
    for (int t = 0; t < m; ++t) {
      const T dx = xi - sx[t], dy = yi - sy[t], dz = zi_ - sz[t];
      const T r2 = dx * dx + dy * dy + dz * dz;
      if (r2 <= P.disp2_sq && base + t != i && sreal[t]) {
        T c0a = T(0), c1a = T(0), c2a = T(0), c3a = T(0);
        const T* aj = sA + t * AVEC;
#pragma unroll
        for (int w = 0; w + 4 <= NFREQ; w += 4) {
          c0a += ai[w] * aj[w];
          c1a += ai[w + 1] * aj[w + 1];
          c2a += ai[w + 2] * aj[w + 2];
          c3a += ai[w + 3] * aj[w + 3];
        }
        c0a += ai[20] * aj[20];
        c1a += ai[21] * aj[21];
        c2a += ai[22] * aj[22];
        const T c6 = (c0a + c1a) + (c2a + c3a);
        const T s2 = si * ss[t];
        const T R0 = P.a1 * s2 + P.a2;
        const T qq = s2 * s2;
        const T r4 = r2 * r2, r6 = r4 * r2, r8 = r4 * r4;
        const T R2 = R0 * R0, R4 = R2 * R2, R6 = R4 * R2, R8 = R4 * R4;
        T F = P.s6 * d4_rcp(r6 + R6) + P.s8 * qq * d4_rcp(r8 + R8);
        if (P.s10k != T(0)) F += P.s10k * qq * qq * d4_rcp(r8 * r2 + R8 * R2);
        acc += c6 * F;
      }
    }
  }
  if (real_i && acc != T(0)) atomicAdd(&A.energy[i], T(-0.5) * acc);
}

// ---------------------------------------------------------- union lists
// CTA per group: sweep all atoms in index order, keep those within the ATM cutoff of
// at least one centre of the group (ordered compaction -> the list stays sorted).
template <typename T>
__global__ void __launch_bounds__(256) large_union(LargeArgs<T> A) {
  __shared__ T cx[GROUP], cy[GROUP], cz[GROUP];
  __shared__ int creal[GROUP];
  __shared__ int warp_tot[8];
  __shared__ int base_s;
  const int g = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < GROUP) {
    const int j = g * GROUP + tid;
    const int zj = j < A.nat ? (int)A.numbers[j] : 0;
    creal[tid] = zj > 0 && zj < NELEM;
    cx[tid] = j < A.nat ? A.pos[3 * j] : T(0);
    cy[tid] = j < A.nat ? A.pos[3 * j + 1] : T(0);
    cz[tid] = j < A.nat ? A.pos[3 * j + 2] : T(0);
  }
  if (tid == 0) base_s = 0;
  __syncthreads();
  int* list = A.ulist + (size_t)g * A.ucap;
  unsigned* masks = A.umask + (size_t)g * A.ucap;
  for (int b0 = 0; b0 < A.nat; b0 += 256) {
    const int x = b0 + tid;
    unsigned m = 0;
    if (x < A.nat) {
      const int zx = (int)A.numbers[x];
      if (zx > 0 && zx < NELEM) {
        const T px = A.pos[3 * x], py = A.pos[3 * x + 1], pz = A.pos[3 * x + 2];
#pragma unroll
        for (int j = 0; j < GROUP; ++j) {
          const T dx = px - cx[j], dy = py - cy[j], dz = pz - cz[j];
          const T r2 = dx * dx + dy * dy + dz * dz;
          if (creal[j] && r2 <= A.par.disp3_sq && x != g * GROUP + j) m |= 1u << j;
        }
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, m != 0);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (m != 0) {
      const int dst = off + __popc(bal & ((1u << lane) - 1u));
      list[dst] = x;
      masks[dst] = m;
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += warp_tot[w];
      base_s += tot;
    }
    __syncthreads();
  }
  if (tid == 0) A.ucount[g] = base_s;
}

// ------------------------------------------------------------------- ATM
// Persistent CTA per centre group.  Prologue: stash (r^2, P, u) of every (centre,
// list atom) pair into an L2-resident per-CTA scratch.  Main loop: tiles X <= Y of the
// union list; a warp owns 4 rows of X, its lanes the 32 columns of Y; the (i,k) stash
// is computed on the fly once per pair and reused for the 16 centres.
template <typename T>
__global__ void __launch_bounds__(256, 2) large_atm(LargeArgs<T> A) {
  __shared__ T cA0[GROUP * AVEC];
  __shared__ T cpx[GROUP], cpy[GROUP], cpz[GROUP], csq[GROUP];
  __shared__ int creal[GROUP];
  // tile buffers: X atoms and Y atoms
  __shared__ T tA0[2][TILE * (AVEC + 1)];
  __shared__ T tpx[2][TILE], tpy[2][TILE], tpz[2][TILE], tsq[2][TILE];
  __shared__ int tidx[2][TILE];
  __shared__ unsigned tmask[2][TILE];
  __shared__ T tst[2][GROUP][3][TILE];  // centre stash of the tile atoms
  __shared__ int gcur;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Par<T>& P = A.par;
  T* const cst = A.cstash + (size_t)blockIdx.x * GROUP * 3 * A.ucap;

  while (true) {
    __syncthreads();
    if (tid == 0) gcur = atomicAdd(A.queue, 1);
    __syncthreads();
    // work item = (centre group, slice): the row tiles tx = slice, slice + nslice, ... of the group
    const int g = A.group_begin + gcur / A.nslice, slice = gcur % A.nslice;
    if (g >= A.group_end) break;
    const int nU = A.ucount[g];
    const int* list = A.ulist + (size_t)g * A.ucap;
    const unsigned* masks = A.umask + (size_t)g * A.ucap;
    // ---- centres
    for (int t = tid; t < GROUP * AVEC; t += 256) {
      const int j = g * GROUP + t / AVEC;
      cA0[t] = j < A.nat ? A.a0[(size_t)j * AVEC + (t % AVEC)] : T(0);
    }
    if (tid < GROUP) {
      const int j = g * GROUP + tid;
      const int zj = j < A.nat ? (int)A.numbers[j] : 0;
      creal[tid] = zj > 0 && zj < NELEM;
      cpx[tid] = j < A.nat ? A.pos[3 * j] : T(0);
      cpy[tid] = j < A.nat ? A.pos[3 * j + 1] : T(0);
      cpz[tid] = j < A.nat ? A.pos[3 * j + 2] : T(0);
      csq[tid] = creal[tid] ? A.tab.sqrt_r4r2[zj] : T(0);
    }
    __syncthreads();
    // ---- prologue: centre stash for the whole list
    for (int t = tid; t < GROUP * nU; t += 256) {
      const int j = t / nU, u = t - j * nU;
      const int x = list[u];
      T a = T(1), Pv = T(0), uv = T(0);
      if (masks[u] >> j & 1u) {
        const T dx = A.pos[3 * x] - cpx[j], dy = A.pos[3 * x + 1] - cpy[j], dz = A.pos[3 * x + 2] - cpz[j];
        const T r2 = dx * dx + dy * dy + dz * dz;
        const T rinv = d4_rcp(d4_sqrt(r2));
        T c6 = T(0);
        const T* ax = A.a0 + (size_t)x * AVEC;
#pragma unroll
        for (int w = 0; w < NFREQ; ++w) c6 += cA0[j * AVEC + w] * ax[w];
        const T R0 = P.a1 * csq[j] * A.tab.sqrt_r4r2[(int)A.numbers[x]] + P.a2;
        a = r2;
        const T ri2 = rinv * rinv;
        Pv = P.fac9 * d4_sqrt(fabs(c6)) * (ri2 * ri2 * rinv);  // P' = P / r^2 (see d4b200_small.cuh)
        uv = d4_zero_damp_arg(R0 * rinv, P.alp3, P.alp16 != 0);
      }
      cst[((size_t)j * 3 + 0) * A.ucap + u] = a;
      cst[((size_t)j * 3 + 1) * A.ucap + u] = Pv;
      cst[((size_t)j * 3 + 2) * A.ucap + u] = uv;
    }
    __syncthreads();

    const int ntile = (nU + TILE - 1) / TILE;
    // stage one tile of the union list: atoms, A0 vectors, centre stash
    auto stage = [&](int side, int tileno) {
      for (int l = tid; l < TILE; l += 256) {
        const int u = tileno * TILE + l;
        const bool ok = u < nU;
        const int x = ok ? list[u] : 0;
        tidx[side][l] = ok ? x : -1;
        tmask[side][l] = ok ? masks[u] : 0u;
        tpx[side][l] = ok ? A.pos[3 * x] : T(0);
        tpy[side][l] = ok ? A.pos[3 * x + 1] : T(0);
        tpz[side][l] = ok ? A.pos[3 * x + 2] : T(0);
        tsq[side][l] = ok ? A.tab.sqrt_r4r2[(int)A.numbers[x]] : T(0);
      }
      for (int r = tid; r < TILE * AVEC; r += 256) {
        const int l = r / AVEC, w = r - l * AVEC;
        const int u = tileno * TILE + l;
        tA0[side][l * (AVEC + 1) + w] = u < nU ? A.a0[(size_t)list[u] * AVEC + w] : T(0);
      }
      for (int r = tid; r < GROUP * 3 * TILE; r += 256) {
        const int jc = r / TILE, l = r - jc * TILE;  // jc = j*3 + component
        const int u = tileno * TILE + l;
        (&tst[side][0][0][0])[jc * TILE + l] = u < nU ? cst[(size_t)jc * A.ucap + u] : T(0);
      }
    };
    for (int tx = slice; tx < ntile; tx += A.nslice) {
      __syncthreads();
      stage(0, tx);  // X rows: once per row tile
      for (int ty = tx; ty < ntile; ++ty) {
        if (ty != tx) __syncthreads();
        stage(1, ty);  // Y columns
        __syncthreads();
        // ---- evaluate: lane = column k of Y, warp rows 4w..4w+3 of X
        const unsigned mk = tmask[1][lane];
        const int kk = tidx[1][lane];
        T colacc = T(0);
        for (int rr = 0; rr < 4; ++rr) {
          const int row = warp * 4 + rr;
          const int ii = tidx[0][row];
          const unsigned mi = tmask[0][row];
          T rowacc = T(0);
          // pair (i,k) valid: both present, distinct, and in diagonal tiles row < column
          const bool pv = ii >= 0 && kk >= 0 && ii != kk && (tx != ty || row < lane) && (mi & mk) != 0u;
          if (pv) {
            const T dx = tpx[0][row] - tpx[1][lane], dy = tpy[0][row] - tpy[1][lane], dz = tpz[0][row] - tpz[1][lane];
            const T c = dx * dx + dy * dy + dz * dz;  // r_ik^2 (no cutoff on this edge)
            const T rinv = d4_rcp(d4_sqrt(c));
            T c6 = T(0);
            const T* ai = &tA0[0][row * (AVEC + 1)];
            const T* ak = &tA0[1][lane * (AVEC + 1)];
#pragma unroll
            for (int w = 0; w < NFREQ; ++w) c6 += ai[w] * ak[w];
            const T R0 = P.a1 * tsq[0][row] * tsq[1][lane] + P.a2;
            const T ri2 = rinv * rinv, c2 = c * c;
            const T Pik = P.fac9 * d4_sqrt(fabs(c6)) * (ri2 * ri2 * rinv);
            const T uik = d4_zero_damp_arg(R0 * rinv, P.alp3, P.alp16 != 0);
            // all centres of the group, branch-free: where i or k is not a neighbour of centre j
            // the stash holds (r^2, P', u) = (1, 0, 0) and the term is exactly zero
            T esum = T(0);
#pragma unroll 8
            for (int j = 0; j < GROUP; ++j) {
              const T a = tst[0][j][0][row], b = tst[1][j][0][lane];  // r_ji^2, r_jk^2
              const T t1 = a - b, t2 = a + b;
              const T s = fma(-t1, t1, c2) * (t2 - c);
              const T abc = (a * b) * c;
              const T t = tst[0][j][2][row] * tst[1][j][2][lane] * uik;
              const T pp = tst[0][j][1][row] * tst[1][j][1][lane] * Pik;
              esum += (pp * fma(T(0.375), s, abc)) * d4_rcp(fma(T(6), t, T(1)));
            }
            rowacc = esum;
            colacc += esum;
          }
          rowacc = warp_sum(rowacc);
          if (lane == 0 && rowacc != T(0)) atomicAdd(&A.energy[ii], rowacc);
        }
        if (colacc != T(0)) atomicAdd(&A.energy[kk], colacc);
      }
    }
  }
}


// ============================ gradient kernels ==============================
// dL/dR, dL/dcn, dL/dq for L = sum_i g_i E_i (tests/kernel_model.py has the algebra).
// Per-pair derivative scalars are 23-term dots with the derivative vectors
// dA/dcn, dA/dq; everything is accumulated per ATOM (row owner / lane owner / centre
// owner), so no per-pair storage is needed.

// Two-body, one thread per row atom i: complete force on i from the pair terms and the
// row's share of dL/dcn_i, dL/dq_i.
template <typename T>
__global__ void __launch_bounds__(128) large_twobody_grad(LargeArgs<T> A, int ncolchunks) {
  extern __shared__ __align__(16) unsigned char dsm[];
  T* const rv = reinterpret_cast<T*>(dsm);           // [3][NFREQ][128] row vectors, transposed
  T* const sA = rv + 3 * NFREQ * 128;               // [64][AVEC] column tile
  T* const sx = sA + 64 * AVEC;
  T* const sy = sx + 64;
  T* const sz = sy + 64;
  T* const ss = sz + 64;
  T* const sg = ss + 64;
  int* const sreal = reinterpret_cast<int*>(sg + 64);
  const int tid = threadIdx.x;
  const int i = A.row_begin + blockIdx.x * 128 + tid;
  const bool act = i < A.row_end;
  const int zi = act ? (int)A.numbers[i] : 0;
  const bool real_i = act && zi > 0 && zi < NELEM;
  for (int w = 0; w < NFREQ; ++w) {
    rv[(0 * NFREQ + w) * 128 + tid] = real_i ? A.aq[(size_t)i * AVEC + w] : T(0);
    rv[(1 * NFREQ + w) * 128 + tid] = real_i ? A.daq_cn[(size_t)i * AVEC + w] : T(0);
    rv[(2 * NFREQ + w) * 128 + tid] = real_i ? A.daq_q[(size_t)i * AVEC + w] : T(0);
  }
  const T xi = act ? A.pos[3 * i] : T(0), yi = act ? A.pos[3 * i + 1] : T(0), zi_ = act ? A.pos[3 * i + 2] : T(0);
  const T si = real_i ? A.tab.sqrt_r4r2[zi] : T(0);
  const T gi = real_i ? (A.gin ? A.gin[i] : T(1)) : T(0);
  const Par<T>& P = A.par;
  const int chunk = (A.nat + ncolchunks - 1) / ncolchunks;
  const int c0 = blockIdx.y * chunk, c1 = min(A.nat, c0 + chunk);
  T fx = T(0), fy = T(0), fz = T(0), dcn = T(0), dq = T(0), e2 = T(0);
  for (int base = c0; base < c1; base += 64) {
    __syncthreads();
    for (int t = tid; t < 64 * AVEC; t += 128) {
      const int j = base + t / AVEC;
      sA[t] = j < c1 ? A.aq[(size_t)j * AVEC + (t % AVEC)] : T(0);
    }
    if (tid < 64) {
      const int j = base + tid;
      const int zj = j < c1 ? (int)A.numbers[j] : 0;
      const bool rj = zj > 0 && zj < NELEM;
      sx[tid] = j < c1 ? A.pos[3 * j] : T(0);
      sy[tid] = j < c1 ? A.pos[3 * j + 1] : T(0);
      sz[tid] = j < c1 ? A.pos[3 * j + 2] : T(0);
      ss[tid] = rj ? A.tab.sqrt_r4r2[zj] : T(0);
      sg[tid] = rj ? (A.gin ? A.gin[j] : T(1)) : T(0);
      sreal[tid] = rj;
    }
    __syncthreads();
    if (!real_i) continue;
    const int m = min(64, c1 - base);
    for (int t = 0; t < m; ++t) {
      const T dx = xi - sx[t], dy = yi - sy[t], dz = zi_ - sz[t];
      const T r2 = dx * dx + dy * dy + dz * dz;
      if (r2 <= P.disp2_sq && base + t != i && sreal[t]) {
        T c6 = T(0), dc = T(0), dqv = T(0);
        const T* aj = sA + t * AVEC;
#pragma unroll
        for (int w = 0; w < NFREQ; ++w) {
          const T v = aj[w];
          c6 += rv[(0 * NFREQ + w) * 128 + tid] * v;
          dc += rv[(1 * NFREQ + w) * 128 + tid] * v;
          dqv += rv[(2 * NFREQ + w) * 128 + tid] * v;
        }
        const T s2 = si * ss[t];
        const T R0 = P.a1 * s2 + P.a2;
        const T qq = s2 * s2;
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
        const T G2 = T(-0.5) * (gi + sg[t]);
        e2 += c6 * F;  // fused energy + gradient call: the row's two-body energy
        const T fc = G2 * c6 * dF;
        fx += fc * dx;
        fy += fc * dy;
        fz += fc * dz;
        dcn += G2 * F * dc;
        dq += G2 * F * dqv;
      }
    }
  }
  if (real_i) {
    atomicAdd(&A.force[3 * i], fx);
    atomicAdd(&A.force[3 * i + 1], fy);
    atomicAdd(&A.force[3 * i + 2], fz);
    atomicAdd(&A.dcn[i], dcn);
    atomicAdd(&A.dq[i], dq);
    if (A.energy && e2 != T(0)) atomicAdd(&A.energy[i], T(-0.5) * e2);
  }
}

// CN chain rule (stage 2, after dL/dcn has been summed over the ranks):
// F_i += sum_j (dL/dcn_i + dL/dcn_j) dcn_ij/dr (R_i - R_j)/r  for the rows of this rank.
template <typename T>
__global__ void __launch_bounds__(128) large_cn_chain(LargeArgs<T> A, int ncolchunks) {
  __shared__ T sx[128], sy[128], sz[128], sr[128], sd[128];
  __shared__ int szn[128];
  const int i = A.row_begin + blockIdx.x * 128 + threadIdx.x;
  const bool act = i < A.row_end;
  const int zraw = act ? (int)A.numbers[i] : 0;
  const int zi = zraw > 0 && zraw < NELEM ? zraw : 0;
  const T xi = act ? A.pos[3 * i] : T(0), yi = act ? A.pos[3 * i + 1] : T(0), zi_ = act ? A.pos[3 * i + 2] : T(0);
  const T ri = zi ? A.tab.rcov[zi] : T(1);
  const T di = zi ? A.dcn[i] : T(0);
  const int chunk = (A.nat + ncolchunks - 1) / ncolchunks;
  const int c0 = blockIdx.y * chunk, c1 = min(A.nat, c0 + chunk);
  T fx = T(0), fy = T(0), fz = T(0);
  for (int base = c0; base < c1; base += 128) {
    const int j = base + threadIdx.x;
    __syncthreads();
    if (j < c1) {
      const int zj = (int)A.numbers[j];
      const bool rj = zj > 0 && zj < NELEM;
      sx[threadIdx.x] = A.pos[3 * j];
      sy[threadIdx.x] = A.pos[3 * j + 1];
      sz[threadIdx.x] = A.pos[3 * j + 2];
      szn[threadIdx.x] = rj ? zj : 0;
      sr[threadIdx.x] = rj ? A.tab.rcov[zj] : T(1);
      sd[threadIdx.x] = rj ? A.dcn[j] : T(0);
    }
    __syncthreads();
    const int m = min(128, c1 - base);
    if (zi) {
      for (int t = 0; t < m; ++t) {
        const T dx = xi - sx[t], dy = yi - sy[t], dz = zi_ - sz[t];
        const T r2 = dx * dx + dy * dy + dz * dz;
        if (r2 <= A.par.cn_sq && base + t != i && szn[t] != 0) {
          const T r = d4_sqrt(r2);
          const T r0inv = d4_rcp(ri + sr[t]);
          const T xx = T(7.5) * (r * r0inv - T(1));
          if (fabs(xx) < T(8.7)) {
            const T dcn = -A.tab.den[zi * NELEM + szn[t]] * T(7.5) * T(0.5641895835477563) * r0inv * d4_exp(-xx * xx);
            const T c = (di + sd[t]) * dcn * d4_rcp(r);
            fx += c * dx;
            fy += c * dy;
            fz += c * dz;
          }
        }
      }
    }
  }
  if (zi) {
    atomicAdd(&A.force[3 * i], fx);
    atomicAdd(&A.force[3 * i + 1], fy);
    atomicAdd(&A.force[3 * i + 2], fz);
  }
}

// ATM gradient, centre groups.  512 threads: warp w owns rows 2w, 2w+1 of the X tile,
// lanes own the columns of the Y tile; loop order centre -> row so that the centre's
// accumulators are four registers reduced once per (tile, centre).
// Four warp-wide sums for the price of two: a transposing butterfly.  After the call the lanes with
// (lane & 7) == 0 hold the totals: lane 0 -> v0, lane 8 -> v1, lane 16 -> v2, lane 24 -> v3 (the other
// lanes hold the same totals of their group of eight).  12 + 6 instead of 40 + 20 shuffle / add instructions.
template <typename T>
__device__ __forceinline__ T warp_sum4(T v0, T v1, T v2, T v3, int lane) {
  const bool h16 = lane & 16, h8 = lane & 8;
  T k0 = h16 ? v2 : v0, k1 = h16 ? v3 : v1;
  k0 += __shfl_xor_sync(0xffffffffu, h16 ? v0 : v2, 16);
  k1 += __shfl_xor_sync(0xffffffffu, h16 ? v1 : v3, 16);
  T k = h8 ? k1 : k0;
  k += __shfl_xor_sync(0xffffffffu, h8 ? k0 : k1, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  return k;
}

constexpr int GSTASH = 5;  // r^2, P, u, D^(j)_jx, D^(x)_xj
template <typename T>
__global__ void __launch_bounds__(512, 1) large_atm_grad(LargeArgs<T> A) {
  extern __shared__ __align__(16) unsigned char dsm[];
  T* p = reinterpret_cast<T*>(dsm);
  T* const cA0 = p;                 p += GROUP * AVEC;         // centres: A0
  T* const cdA = p;                 p += GROUP * AVEC;         // centres: dA0/dcn
  T* const cpx = p;                 p += GROUP;
  T* const cpy = p;                 p += GROUP;
  T* const cpz = p;                 p += GROUP;
  T* const csq = p;                 p += GROUP;
  T* const cg = p;                  p += GROUP;
  T* const tA0 = p;                 p += 2 * TILE * (AVEC + 1);  // tile atoms: A0
  T* const tdA = p;                 p += 2 * TILE * (AVEC + 1);  // tile atoms: dA0/dcn
  T* const tpx = p;                 p += 2 * TILE;
  T* const tpy = p;                 p += 2 * TILE;
  T* const tpz = p;                 p += 2 * TILE;
  T* const tsq = p;                 p += 2 * TILE;
  T* const tg = p;                  p += 2 * TILE;
  T* const tst = p;                 p += 2 * GROUP * GSTASH * TILE;  // [side][j][comp][l]
  int* const tidx = reinterpret_cast<int*>(p);
  unsigned* const tmask = reinterpret_cast<unsigned*>(tidx + 2 * TILE);
  int* const creal = reinterpret_cast<int*>(tmask + 2 * TILE);
  int* const gcur = creal + GROUP;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Par<T>& P = A.par;
  T* const cst = A.cstash + (size_t)blockIdx.x * GROUP * GSTASH * A.ucap;
#define TST(side, j, comp, l) tst[(((side) * GROUP + (j)) * GSTASH + (comp)) * TILE + (l)]

  while (true) {
    __syncthreads();
    if (tid == 0) *gcur = atomicAdd(A.queue, 1);
    __syncthreads();
    const int g = A.group_begin + *gcur / A.nslice, slice = *gcur % A.nslice;
    if (g >= A.group_end) break;
    const int nU = A.ucount[g];
    const int* list = A.ulist + (size_t)g * A.ucap;
    const unsigned* masks = A.umask + (size_t)g * A.ucap;
    for (int t = tid; t < GROUP * AVEC; t += 512) {
      const int j = g * GROUP + t / AVEC;
      cA0[t] = j < A.nat ? A.a0[(size_t)j * AVEC + (t % AVEC)] : T(0);
      cdA[t] = j < A.nat ? A.da0_cn[(size_t)j * AVEC + (t % AVEC)] : T(0);
    }
    if (tid < GROUP) {
      const int j = g * GROUP + tid;
      const int zj = j < A.nat ? (int)A.numbers[j] : 0;
      creal[tid] = zj > 0 && zj < NELEM;
      cpx[tid] = j < A.nat ? A.pos[3 * j] : T(0);
      cpy[tid] = j < A.nat ? A.pos[3 * j + 1] : T(0);
      cpz[tid] = j < A.nat ? A.pos[3 * j + 2] : T(0);
      csq[tid] = creal[tid] ? A.tab.sqrt_r4r2[zj] : T(0);
      cg[tid] = creal[tid] ? (A.gin ? A.gin[j] : T(1)) : T(0);
    }
    __syncthreads();
    // ---- prologue: centre stash (5 values) for the whole list
    for (int t = tid; t < GROUP * nU; t += 512) {
      const int j = t / nU, u = t - j * nU;
      const int x = list[u];
      T a = T(1), Pv = T(0), uv = T(0), Dj = T(0), Dx = T(0);
      if (masks[u] >> j & 1u) {
        const T dx = A.pos[3 * x] - cpx[j], dy = A.pos[3 * x + 1] - cpy[j], dz = A.pos[3 * x + 2] - cpz[j];
        const T r2 = dx * dx + dy * dy + dz * dz;
        const T rinv = d4_rcp(d4_sqrt(r2));
        T c6 = T(0), dj = T(0), dxv = T(0);
        const T* ax = A.a0 + (size_t)x * AVEC;
        const T* dax = A.da0_cn + (size_t)x * AVEC;
#pragma unroll
        for (int w = 0; w < NFREQ; ++w) {
          c6 += cA0[j * AVEC + w] * ax[w];
          dj += cdA[j * AVEC + w] * ax[w];
          dxv += dax[w] * cA0[j * AVEC + w];
        }
        const T R0 = P.a1 * csq[j] * A.tab.sqrt_r4r2[(int)A.numbers[x]] + P.a2;
        a = r2;
        Pv = P.fac9 * d4_sqrt(fabs(c6)) * (rinv * rinv * rinv);
        uv = d4_zero_damp_arg(R0 * rinv, P.alp3, P.alp16 != 0);
        const T h = c6 != T(0) ? T(0.5) * d4_rcp(c6) : T(0);
        Dj = dj * h;
        Dx = dxv * h;
      }
      cst[((size_t)j * GSTASH + 0) * A.ucap + u] = a;
      cst[((size_t)j * GSTASH + 1) * A.ucap + u] = Pv;
      cst[((size_t)j * GSTASH + 2) * A.ucap + u] = uv;
      cst[((size_t)j * GSTASH + 3) * A.ucap + u] = Dj;
      cst[((size_t)j * GSTASH + 4) * A.ucap + u] = Dx;
    }
    __syncthreads();

    const int ntile = (nU + TILE - 1) / TILE;
    for (int tx = slice; tx < ntile; tx += A.nslice) {
      for (int ty = tx; ty < ntile; ++ty) {
        __syncthreads();
        for (int t = tid; t < 2 * TILE; t += 512) {
          const int side = t / TILE, l = t - side * TILE;
          const int u = (side == 0 ? tx : ty) * TILE + l;
          const bool ok = u < nU;
          const int x = ok ? list[u] : 0;
          tidx[t] = ok ? x : -1;
          tmask[t] = ok ? masks[u] : 0u;
          tpx[t] = ok ? A.pos[3 * x] : T(0);
          tpy[t] = ok ? A.pos[3 * x + 1] : T(0);
          tpz[t] = ok ? A.pos[3 * x + 2] : T(0);
          tsq[t] = ok ? A.tab.sqrt_r4r2[(int)A.numbers[x]] : T(0);
          tg[t] = ok ? (A.gin ? A.gin[x] : T(1)) : T(0);
        }
        for (int t = tid; t < 2 * TILE * AVEC; t += 512) {
          const int side = t / (TILE * AVEC), r = t - side * TILE * AVEC;
          const int l = r / AVEC, w = r - l * AVEC;
          const int u = (side == 0 ? tx : ty) * TILE + l;
          const size_t src = u < nU ? (size_t)list[u] * AVEC + w : 0;
          tA0[(side * TILE + l) * (AVEC + 1) + w] = u < nU ? A.a0[src] : T(0);
          tdA[(side * TILE + l) * (AVEC + 1) + w] = u < nU ? A.da0_cn[src] : T(0);
        }
        for (int t = tid; t < 2 * GROUP * GSTASH * TILE; t += 512) {
          const int side = t / (GROUP * GSTASH * TILE), r = t - side * GROUP * GSTASH * TILE;
          const int jc = r / TILE, l = r - jc * TILE;
          const int u = (side == 0 ? tx : ty) * TILE + l;
          tst[t] = u < nU ? cst[(size_t)jc * A.ucap + u] : T(0);
        }
        __syncthreads();
        // ---- faces of this lane: rows 2w, 2w+1 x column `lane`
        const int kk = tidx[TILE + lane];
        const unsigned mk = tmask[TILE + lane];
        const T gk = tg[TILE + lane];
        const T kx = tpx[TILE + lane], ky = tpy[TILE + lane], kz = tpz[TILE + lane];
        T fc_[2], fP[2], fu[2], fDi[2], fDk[2];
        bool pv[2];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int row = warp * 2 + rr;
          const int ii = tidx[row];
          pv[rr] = ii >= 0 && kk >= 0 && ii != kk && (tx != ty || row < lane) && (tmask[row] & mk) != 0u;
          fc_[rr] = T(1), fP[rr] = T(0), fu[rr] = T(0), fDi[rr] = T(0), fDk[rr] = T(0);
          const T rx = tpx[row] - kx, ry = tpy[row] - ky, rz = tpz[row] - kz;  // R_i - R_k
          if (pv[rr]) {
            const T c = rx * rx + ry * ry + rz * rz;
            const T rinv = d4_rcp(d4_sqrt(c));
            T c6 = T(0), di = T(0), dk = T(0);
            const T* ai = &tA0[row * (AVEC + 1)];
            const T* ak = &tA0[(TILE + lane) * (AVEC + 1)];
            const T* dai = &tdA[row * (AVEC + 1)];
            const T* dak = &tdA[(TILE + lane) * (AVEC + 1)];
#pragma unroll
            for (int w = 0; w < NFREQ; ++w) {
              c6 += ai[w] * ak[w];
              di += dai[w] * ak[w];
              dk += dak[w] * ai[w];
            }
            const T R0 = P.a1 * tsq[row] * tsq[TILE + lane] + P.a2;
            const T h = c6 != T(0) ? T(0.5) * d4_rcp(c6) : T(0);
            fc_[rr] = c;
            fP[rr] = P.fac9 * d4_sqrt(fabs(c6)) * (rinv * rinv * rinv);
            fu[rr] = d4_zero_damp_arg(R0 * rinv, P.alp3, P.alp16 != 0);
            fDi[rr] = di * h;
            fDk[rr] = dk * h;
          }
        }
        // Forces without per-visit difference vectors: with the edge derivatives da (j,i), db (j,k), dc (i,k)
        //   F_i = R_i sum_j da - sum_j da R_j + (R_i - R_k) sum_j dc      per (row, lane)
        //   F_k = R_k sum db - sum db R_j - sum_rows (R_i - R_k) sum_j dc  per lane
        //   F_j = R_j sum (da + db) - sum (da R_i + db R_k)               per centre
        // i.e. 17 instead of 30 FP64 operations per visit for the force part.
        T ksb = T(0), kbx = T(0), kby = T(0), kbz = T(0), kdc = T(0);   // column atom k
        T isa[2] = {T(0), T(0)}, iax[2] = {T(0), T(0)}, iay[2] = {T(0), T(0)}, iaz[2] = {T(0), T(0)};
        T isc[2] = {T(0), T(0)}, idc[2] = {T(0), T(0)};
        T ie[2] = {T(0), T(0)}, ke = T(0);  // fused energy + gradient call: E_i += e, E_k += e per centre (large_atm)
        for (int j = 0; j < GROUP; ++j) {
          T jvx = T(0), jvy = T(0), jvz = T(0), js = T(0), jdc = T(0);   // centre j
          const T b = TST(1, j, 0, lane), Pb = TST(1, j, 1, lane), ub = TST(1, j, 2, lane);
          const T Djk = TST(1, j, 3, lane), Dkj = TST(1, j, 4, lane);
          const T cx = cpx[j], cy = cpy[j], cz = cpz[j];
          const bool kin = mk >> j & 1u;
          // both rows of the lane as ONE straight-line block (masked instead of branched): two independent
          // dependency chains in flight
          const bool in0 = tmask[warp * 2] >> j & 1u, in1 = tmask[warp * 2 + 1] >> j & 1u;  // warp-uniform
          if (in0 | in1) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const int row = warp * 2 + rr;
              const bool act = pv[rr] && kin && (rr ? in1 : in0);
              // masked lanes run on a unit triangle (finite everywhere) and contribute psf = 0
              const T a = act ? TST(0, j, 0, row) : T(1), c = act ? fc_[rr] : T(1), bm = act ? b : T(1);
              const T X = a + bm - c, Y = a - bm + c, Z = bm + c - a;
              const T s = X * Y * Z;
              const T abc = a * bm * c;
              const T t = TST(0, j, 2, row) * ub * fu[rr];
              const T d = T(1) + T(6) * t;
              const T inv = d4_rcp(abc * d);
              const T Q = inv * d, f = inv * abc;
              const T psf = act ? TST(0, j, 1, row) * Pb * fP[rr] * f : T(0);
              const T e = (T(0.375) * s * Q + T(1)) * psf;
              const T W = tg[row] + gk;
              const T common = e * (T(-2.5) + T(3) * P.alp3 * f * t) + psf;
              const T k3 = T(0.375) * psf * Q;
              const T yz = Y * Z, xz = X * Z, xy = X * Y;
              const T da = T(2) * W * (common * (Q * bm * c) + k3 * (yz + xz - xy));   // (j,i)
              const T db = T(2) * W * (common * (Q * a * c) + k3 * (yz - xz + xy));   // (j,k)
              const T dc = T(2) * W * (common * (Q * a * bm) + k3 * (xz + xy - yz));   // (i,k)
              isa[rr] += da;
              iax[rr] = fma(da, cx, iax[rr]), iay[rr] = fma(da, cy, iay[rr]), iaz[rr] = fma(da, cz, iaz[rr]);
              isc[rr] += dc;
              ksb += db;
              kbx = fma(db, cx, kbx), kby = fma(db, cy, kby), kbz = fma(db, cz, kbz);
              jvx = fma(da, tpx[row], fma(db, kx, jvx));
              jvy = fma(da, tpy[row], fma(db, ky, jvy));
              jvz = fma(da, tpz[row], fma(db, kz, jvz));
              js += da + db;
              const T We = W * e;
              ie[rr] += e;
              ke += e;
              idc[rr] += We * (TST(0, j, 4, row) + fDi[rr]);
              kdc += We * (Dkj + fDk[rr]);
              jdc += We * (TST(0, j, 3, row) + Djk);
            }
          }
          {  // x, y, z, dcn of the centre end up in the lanes 0, 8, 16, 24
            const T tot = warp_sum4(fma(js, cx, -jvx), fma(js, cy, -jvy), fma(js, cz, -jvz), jdc, lane);
            if ((lane & 7) == 0 && tot != T(0)) {
              const int jj = g * GROUP + j, which = lane >> 3;
              atomicAdd(which < 3 ? &A.force[3 * jj + which] : &A.dcn[jj], tot);
            }
          }
        }
        T kfx = fma(ksb, kx, -kbx), kfy = fma(ksb, ky, -kby), kfz = fma(ksb, kz, -kbz);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int row = warp * 2 + rr;
          const T px = tpx[row], py = tpy[row], pz = tpz[row];
          const T cxk = isc[rr] * (px - kx), cyk = isc[rr] * (py - ky), czk = isc[rr] * (pz - kz);  // (i,k) edge
          kfx -= cxk, kfy -= cyk, kfz -= czk;
          const T tot = warp_sum4(fma(isa[rr], px, -iax[rr]) + cxk, fma(isa[rr], py, -iay[rr]) + cyk,
                                  fma(isa[rr], pz, -iaz[rr]) + czk, idc[rr], lane);
          const int ii = tidx[row];
          if ((lane & 7) == 0 && ii >= 0 && tot != T(0)) {
            const int which = lane >> 3;
            atomicAdd(which < 3 ? &A.force[3 * ii + which] : &A.dcn[ii], tot);
          }
          if (A.energy) {
            const T se_ = warp_sum(ie[rr]);
            if (lane == 0 && ii >= 0 && se_ != T(0)) atomicAdd(&A.energy[ii], se_);
          }
        }
        if (A.energy && kk >= 0 && ke != T(0)) atomicAdd(&A.energy[kk], ke);
        if (kk >= 0 && (kfx != T(0) || kfy != T(0) || kfz != T(0) || kdc != T(0))) {
          atomicAdd(&A.force[3 * kk], kfx);
          atomicAdd(&A.force[3 * kk + 1], kfy);
          atomicAdd(&A.force[3 * kk + 2], kfz);
          atomicAdd(&A.dcn[kk], kdc);
        }
      }
    }
  }
#undef TST
}

// The same kernel as a double-buffered pipeline: everything a tile pair needs sits in per-CTA global scratch in
// the layout of the shared-memory arrays (written once per centre group by the prologue), and one thread moves
// the NEXT pair into the other buffer with bulk-async copies (cp.async.bulk + mbarrier, UBLKCP in SASS) while
// the CTA evaluates the current one: one block barrier per pair, no load latency on the path.
template <typename T>
__host__ __device__ constexpr size_t atm_pipe_buffer_bytes() {
  return sizeof(T) * (2 * 2 * TILE * (AVEC + 1) + 5 * 2 * TILE + 2 * GROUP * GSTASH * TILE) + sizeof(int) * 4 * TILE;
}
template <typename T>
__host__ __device__ constexpr size_t atm_pipe_smem() {
  return sizeof(T) * (2 * GROUP * AVEC + 5 * GROUP) + 2 * atm_pipe_buffer_bytes<T>() + sizeof(int) * (GROUP + 4) + 32;
}
// per-CTA scratch of the pipeline, in tiles of TILE list atoms
template <typename T>
__host__ __device__ constexpr size_t atm_pipe_tile_bytes() {
  return sizeof(T) * (GROUP * GSTASH * TILE + 2 * TILE * (AVEC + 1) + 5 * TILE) + sizeof(int) * 2 * TILE;
}

template <typename T>
__global__ void __launch_bounds__(512, 1) large_atm_grad_pipe(LargeArgs<T> A) {
  extern __shared__ __align__(16) unsigned char dsm[];
  T* p = reinterpret_cast<T*>(dsm);
  T* const cA0 = p;                 p += GROUP * AVEC;         // centres: A0
  T* const cdA = p;                 p += GROUP * AVEC;         // centres: dA0/dcn
  T* const cpx = p;                 p += GROUP;
  T* const cpy = p;                 p += GROUP;
  T* const cpz = p;                 p += GROUP;
  T* const csq = p;                 p += GROUP;
  T* const cg = p;                  p += GROUP;
  unsigned char* const buf0 = reinterpret_cast<unsigned char*>(p);
  int* const creal = reinterpret_cast<int*>(buf0 + 2 * atm_pipe_buffer_bytes<T>());
  int* const gcur = creal + GROUP;
  unsigned long long* const bars = reinterpret_cast<unsigned long long*>(
      (reinterpret_cast<uintptr_t>(gcur + 4) + 7) & ~uintptr_t(7));
  // arrays of the CURRENT buffer (same names and layouts as in large_atm_grad)
  T *tA0, *tdA, *tpx, *tpy, *tpz, *tsq, *tg, *tst;
  int* tidx;
  unsigned* tmask;
  auto select = [&](int b) {
    T* q = reinterpret_cast<T*>(buf0 + (size_t)b * atm_pipe_buffer_bytes<T>());
    tA0 = q, q += 2 * TILE * (AVEC + 1);
    tdA = q, q += 2 * TILE * (AVEC + 1);
    tpx = q, q += 2 * TILE;
    tpy = q, q += 2 * TILE;
    tpz = q, q += 2 * TILE;
    tsq = q, q += 2 * TILE;
    tg = q, q += 2 * TILE;
    tst = q, q += 2 * GROUP * GSTASH * TILE;
    tidx = reinterpret_cast<int*>(q);
    tmask = reinterpret_cast<unsigned*>(tidx + 2 * TILE);
  };
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Par<T>& P = A.par;
  // scratch of this CTA: [tile] x { stash [GROUP][GSTASH][TILE], A0 [TILE][AVEC+1], dA0 [TILE][AVEC+1],
  // x, y, z, sqrt_r4r2, g [TILE] each, index [TILE], mask [TILE] }
  const int ntmax = (A.ucap + TILE - 1) / TILE;
  unsigned char* const scr = reinterpret_cast<unsigned char*>(A.cstash) + (size_t)blockIdx.x * ntmax * atm_pipe_tile_bytes<T>();
  constexpr size_t O_A0 = sizeof(T) * GROUP * GSTASH * TILE, O_DA = O_A0 + sizeof(T) * TILE * (AVEC + 1),
                   O_AT = O_DA + sizeof(T) * TILE * (AVEC + 1), O_IX = O_AT + sizeof(T) * 5 * TILE,
                   O_MK = O_IX + sizeof(int) * TILE;
  auto tile_ptr = [&](int tile) { return scr + (size_t)tile * atm_pipe_tile_bytes<T>(); };
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
  }
  unsigned parity0 = 0, parity1 = 0;
  // thread 0: bring tile pair (tx, ty) into buffer b
  auto issue = [&](int b, int tx, int ty) {
    unsigned char* dst = buf0 + (size_t)b * atm_pipe_buffer_bytes<T>();
    T* q = reinterpret_cast<T*>(dst);
    T* const dA0 = q;  q += 2 * TILE * (AVEC + 1);
    T* const ddA = q;  q += 2 * TILE * (AVEC + 1);
    T* const dat = q;  q += 5 * 2 * TILE;
    T* const dst_st = q;  q += 2 * GROUP * GSTASH * TILE;
    int* const dix = reinterpret_cast<int*>(q);
    unsigned* const dmk = reinterpret_cast<unsigned*>(dix + 2 * TILE);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads of this buffer are done
    mbar_expect_tx(&bars[b], (unsigned)(2 * atm_pipe_tile_bytes<T>()));
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const unsigned char* src = tile_ptr(side == 0 ? tx : ty);
      bulk_g2s(dst_st + side * GROUP * GSTASH * TILE, src, sizeof(T) * GROUP * GSTASH * TILE, &bars[b]);
      bulk_g2s(dA0 + side * TILE * (AVEC + 1), src + O_A0, sizeof(T) * TILE * (AVEC + 1), &bars[b]);
      bulk_g2s(ddA + side * TILE * (AVEC + 1), src + O_DA, sizeof(T) * TILE * (AVEC + 1), &bars[b]);
#pragma unroll
      for (int c = 0; c < 5; ++c)
        bulk_g2s(dat + (c * 2 + side) * TILE, src + O_AT + sizeof(T) * c * TILE, sizeof(T) * TILE, &bars[b]);
      bulk_g2s(dix + side * TILE, src + O_IX, sizeof(int) * TILE, &bars[b]);
      bulk_g2s(dmk + side * TILE, src + O_MK, sizeof(unsigned) * TILE, &bars[b]);
    }
  };
#define TST(side, j, comp, l) tst[(((side) * GROUP + (j)) * GSTASH + (comp)) * TILE + (l)]

  while (true) {
    __syncthreads();
    if (tid == 0) *gcur = atomicAdd(A.queue, 1);
    __syncthreads();
    const int g = A.group_begin + *gcur / A.nslice, slice = *gcur % A.nslice;
    if (g >= A.group_end) break;
    const int nU = A.ucount[g];
    const int* list = A.ulist + (size_t)g * A.ucap;
    const unsigned* masks = A.umask + (size_t)g * A.ucap;
    for (int t = tid; t < GROUP * AVEC; t += 512) {
      const int j = g * GROUP + t / AVEC;
      cA0[t] = j < A.nat ? A.a0[(size_t)j * AVEC + (t % AVEC)] : T(0);
      cdA[t] = j < A.nat ? A.da0_cn[(size_t)j * AVEC + (t % AVEC)] : T(0);
    }
    if (tid < GROUP) {
      const int j = g * GROUP + tid;
      const int zj = j < A.nat ? (int)A.numbers[j] : 0;
      creal[tid] = zj > 0 && zj < NELEM;
      cpx[tid] = j < A.nat ? A.pos[3 * j] : T(0);
      cpy[tid] = j < A.nat ? A.pos[3 * j + 1] : T(0);
      cpz[tid] = j < A.nat ? A.pos[3 * j + 2] : T(0);
      csq[tid] = creal[tid] ? A.tab.sqrt_r4r2[zj] : T(0);
      cg[tid] = creal[tid] ? (A.gin ? A.gin[j] : T(1)) : T(0);
    }
    __syncthreads();
    // ---- prologue: centre stash (5 values) for the whole list
    for (int t = tid; t < GROUP * nU; t += 512) {
      const int j = t / nU, u = t - j * nU;
      const int x = list[u];
      T a = T(1), Pv = T(0), uv = T(0), Dj = T(0), Dx = T(0);
      if (masks[u] >> j & 1u) {
        const T dx = A.pos[3 * x] - cpx[j], dy = A.pos[3 * x + 1] - cpy[j], dz = A.pos[3 * x + 2] - cpz[j];
        const T r2 = dx * dx + dy * dy + dz * dz;
        const T rinv = d4_rcp(d4_sqrt(r2));
        T c6 = T(0), dj = T(0), dxv = T(0);
        const T* ax = A.a0 + (size_t)x * AVEC;
        const T* dax = A.da0_cn + (size_t)x * AVEC;
#pragma unroll
        for (int w = 0; w < NFREQ; ++w) {
          c6 += cA0[j * AVEC + w] * ax[w];
          dj += cdA[j * AVEC + w] * ax[w];
          dxv += dax[w] * cA0[j * AVEC + w];
        }
        const T R0 = P.a1 * csq[j] * A.tab.sqrt_r4r2[(int)A.numbers[x]] + P.a2;
        a = r2;
        Pv = P.fac9 * d4_sqrt(fabs(c6)) * (rinv * rinv * rinv);
        uv = d4_zero_damp_arg(R0 * rinv, P.alp3, P.alp16 != 0);
        const T h = c6 != T(0) ? T(0.5) * d4_rcp(c6) : T(0);
        Dj = dj * h;
        Dx = dxv * h;
      }
      {
        T* const st = reinterpret_cast<T*>(tile_ptr(u / TILE)) + (size_t)j * GSTASH * TILE + (u % TILE);
        st[0 * TILE] = a;
        st[1 * TILE] = Pv;
        st[2 * TILE] = uv;
        st[3 * TILE] = Dj;
        st[4 * TILE] = Dx;
      }
    }
    const int ntile = (nU + TILE - 1) / TILE;
    // padding of the last tile, and the tile atoms themselves (what large_atm_grad gathers per pair)
    for (int t = tid; t < GROUP * GSTASH * (ntile * TILE - nU); t += 512) {
      const int jc = t / (ntile * TILE - nU), u = nU + t % (ntile * TILE - nU);
      reinterpret_cast<T*>(tile_ptr(u / TILE))[(size_t)jc * TILE + (u % TILE)] = T(0);
    }
    for (int u = tid; u < ntile * TILE; u += 512) {
      unsigned char* tp = tile_ptr(u / TILE);
      const int l = u % TILE;
      const bool ok = u < nU;
      const int x = ok ? list[u] : 0;
      T* at = reinterpret_cast<T*>(tp + O_AT);
      at[0 * TILE + l] = ok ? A.pos[3 * x] : T(0);
      at[1 * TILE + l] = ok ? A.pos[3 * x + 1] : T(0);
      at[2 * TILE + l] = ok ? A.pos[3 * x + 2] : T(0);
      at[3 * TILE + l] = ok ? A.tab.sqrt_r4r2[(int)A.numbers[x]] : T(0);
      at[4 * TILE + l] = ok ? (A.gin ? A.gin[x] : T(1)) : T(0);
      reinterpret_cast<int*>(tp + O_IX)[l] = ok ? x : -1;
      reinterpret_cast<unsigned*>(tp + O_MK)[l] = ok ? masks[u] : 0u;
    }
    for (int t = tid; t < ntile * TILE * (AVEC + 1); t += 512) {
      const int u = t / (AVEC + 1), w = t - u * (AVEC + 1);
      const bool ok = u < nU && w < AVEC;
      const size_t src = ok ? (size_t)list[u] * AVEC + w : 0;
      unsigned char* tp = tile_ptr(u / TILE);
      reinterpret_cast<T*>(tp + O_A0)[(u % TILE) * (AVEC + 1) + w] = ok ? A.a0[src] : T(0);
      reinterpret_cast<T*>(tp + O_DA)[(u % TILE) * (AVEC + 1) + w] = ok ? A.da0_cn[src] : T(0);
    }
    asm volatile("fence.proxy.async;" ::: "memory");  // the scratch is read through the async proxy
    __syncthreads();

    // ---- tile pairs (tx <= ty, tx = slice, slice + nslice, ...) as one sequence, one pair in flight ahead
    int tx = slice, ty = slice, nb = 0;
    if (tx < ntile && tid == 0) issue(0, tx, ty);
    while (tx < ntile) {
      int nx = tx, ny = ty + 1;
      if (ny >= ntile) nx = tx + A.nslice, ny = nx;
      if (nx < ntile && tid == 0) issue(nb ^ 1, nx, ny);
      select(nb);
      if (nb == 0) {
        mbar_wait(&bars[0], parity0);
        parity0 ^= 1u;
      } else {
        mbar_wait(&bars[1], parity1);
        parity1 ^= 1u;
      }
      {
        {
        // ---- faces of this lane: rows 2w, 2w+1 x column `lane`
        const int kk = tidx[TILE + lane];
        const unsigned mk = tmask[TILE + lane];
        const T gk = tg[TILE + lane];
        const T kx = tpx[TILE + lane], ky = tpy[TILE + lane], kz = tpz[TILE + lane];
        T fc_[2], fP[2], fu[2], fDi[2], fDk[2];
        bool pv[2];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int row = warp * 2 + rr;
          const int ii = tidx[row];
          pv[rr] = ii >= 0 && kk >= 0 && ii != kk && (tx != ty || row < lane) && (tmask[row] & mk) != 0u;
          fc_[rr] = T(1), fP[rr] = T(0), fu[rr] = T(0), fDi[rr] = T(0), fDk[rr] = T(0);
          const T rx = tpx[row] - kx, ry = tpy[row] - ky, rz = tpz[row] - kz;  // R_i - R_k
          if (pv[rr]) {
            const T c = rx * rx + ry * ry + rz * rz;
            const T rinv = d4_rcp(d4_sqrt(c));
            T c6 = T(0), di = T(0), dk = T(0);
            const T* ai = &tA0[row * (AVEC + 1)];
            const T* ak = &tA0[(TILE + lane) * (AVEC + 1)];
            const T* dai = &tdA[row * (AVEC + 1)];
            const T* dak = &tdA[(TILE + lane) * (AVEC + 1)];
#pragma unroll
            for (int w = 0; w < NFREQ; ++w) {
              c6 += ai[w] * ak[w];
              di += dai[w] * ak[w];
              dk += dak[w] * ai[w];
            }
            const T R0 = P.a1 * tsq[row] * tsq[TILE + lane] + P.a2;
            const T h = c6 != T(0) ? T(0.5) * d4_rcp(c6) : T(0);
            fc_[rr] = c;
            fP[rr] = P.fac9 * d4_sqrt(fabs(c6)) * (rinv * rinv * rinv);
            fu[rr] = d4_zero_damp_arg(R0 * rinv, P.alp3, P.alp16 != 0);
            fDi[rr] = di * h;
            fDk[rr] = dk * h;
          }
        }
        // Forces without per-visit difference vectors: with the edge derivatives da (j,i), db (j,k), dc (i,k)
        //   F_i = R_i sum_j da - sum_j da R_j + (R_i - R_k) sum_j dc      per (row, lane)
        //   F_k = R_k sum db - sum db R_j - sum_rows (R_i - R_k) sum_j dc  per lane
        //   F_j = R_j sum (da + db) - sum (da R_i + db R_k)               per centre
        // i.e. 17 instead of 30 FP64 operations per visit for the force part.
        T ksb = T(0), kbx = T(0), kby = T(0), kbz = T(0), kdc = T(0);   // column atom k
        T isa[2] = {T(0), T(0)}, iax[2] = {T(0), T(0)}, iay[2] = {T(0), T(0)}, iaz[2] = {T(0), T(0)};
        T isc[2] = {T(0), T(0)}, idc[2] = {T(0), T(0)};
        T ie[2] = {T(0), T(0)}, ke = T(0);  // fused energy + gradient call: E_i += e, E_k += e per centre (large_atm)
        for (int j = 0; j < GROUP; ++j) {
          T jvx = T(0), jvy = T(0), jvz = T(0), js = T(0), jdc = T(0);   // centre j
          const T b = TST(1, j, 0, lane), Pb = TST(1, j, 1, lane), ub = TST(1, j, 2, lane);
          const T Djk = TST(1, j, 3, lane), Dkj = TST(1, j, 4, lane);
          const T cx = cpx[j], cy = cpy[j], cz = cpz[j];
          const bool kin = mk >> j & 1u;
          // both rows of the lane as ONE straight-line block (masked instead of branched): two independent
          // dependency chains in flight
          const bool in0 = tmask[warp * 2] >> j & 1u, in1 = tmask[warp * 2 + 1] >> j & 1u;  // warp-uniform
          if (in0 | in1) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const int row = warp * 2 + rr;
              const bool act = pv[rr] && kin && (rr ? in1 : in0);
              // masked lanes run on a unit triangle (finite everywhere) and contribute psf = 0
              const T a = act ? TST(0, j, 0, row) : T(1), c = act ? fc_[rr] : T(1), bm = act ? b : T(1);
              const T X = a + bm - c, Y = a - bm + c, Z = bm + c - a;
              const T s = X * Y * Z;
              const T abc = a * bm * c;
              const T t = TST(0, j, 2, row) * ub * fu[rr];
              const T d = T(1) + T(6) * t;
              const T inv = d4_rcp(abc * d);
              const T Q = inv * d, f = inv * abc;
              const T psf = act ? TST(0, j, 1, row) * Pb * fP[rr] * f : T(0);
              const T e = (T(0.375) * s * Q + T(1)) * psf;
              const T W = tg[row] + gk;
              const T common = e * (T(-2.5) + T(3) * P.alp3 * f * t) + psf;
              const T k3 = T(0.375) * psf * Q;
              const T yz = Y * Z, xz = X * Z, xy = X * Y;
              const T da = T(2) * W * (common * (Q * bm * c) + k3 * (yz + xz - xy));   // (j,i)
              const T db = T(2) * W * (common * (Q * a * c) + k3 * (yz - xz + xy));   // (j,k)
              const T dc = T(2) * W * (common * (Q * a * bm) + k3 * (xz + xy - yz));   // (i,k)
              isa[rr] += da;
              iax[rr] = fma(da, cx, iax[rr]), iay[rr] = fma(da, cy, iay[rr]), iaz[rr] = fma(da, cz, iaz[rr]);
              isc[rr] += dc;
              ksb += db;
              kbx = fma(db, cx, kbx), kby = fma(db, cy, kby), kbz = fma(db, cz, kbz);
              jvx = fma(da, tpx[row], fma(db, kx, jvx));
              jvy = fma(da, tpy[row], fma(db, ky, jvy));
              jvz = fma(da, tpz[row], fma(db, kz, jvz));
              js += da + db;
              const T We = W * e;
              ie[rr] += e;
              ke += e;
              idc[rr] += We * (TST(0, j, 4, row) + fDi[rr]);
              kdc += We * (Dkj + fDk[rr]);
              jdc += We * (TST(0, j, 3, row) + Djk);
            }
          }
          {  // x, y, z, dcn of the centre end up in the lanes 0, 8, 16, 24
            const T tot = warp_sum4(fma(js, cx, -jvx), fma(js, cy, -jvy), fma(js, cz, -jvz), jdc, lane);
            if ((lane & 7) == 0 && tot != T(0)) {
              const int jj = g * GROUP + j, which = lane >> 3;
              atomicAdd(which < 3 ? &A.force[3 * jj + which] : &A.dcn[jj], tot);
            }
          }
        }
        T kfx = fma(ksb, kx, -kbx), kfy = fma(ksb, ky, -kby), kfz = fma(ksb, kz, -kbz);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int row = warp * 2 + rr;
          const T px = tpx[row], py = tpy[row], pz = tpz[row];
          const T cxk = isc[rr] * (px - kx), cyk = isc[rr] * (py - ky), czk = isc[rr] * (pz - kz);  // (i,k) edge
          kfx -= cxk, kfy -= cyk, kfz -= czk;
          const T tot = warp_sum4(fma(isa[rr], px, -iax[rr]) + cxk, fma(isa[rr], py, -iay[rr]) + cyk,
                                  fma(isa[rr], pz, -iaz[rr]) + czk, idc[rr], lane);
          const int ii = tidx[row];
          if ((lane & 7) == 0 && ii >= 0 && tot != T(0)) {
            const int which = lane >> 3;
            atomicAdd(which < 3 ? &A.force[3 * ii + which] : &A.dcn[ii], tot);
          }
          if (A.energy) {
            const T se_ = warp_sum(ie[rr]);
            if (lane == 0 && ii >= 0 && se_ != T(0)) atomicAdd(&A.energy[ii], se_);
          }
        }
        if (A.energy && kk >= 0 && ke != T(0)) atomicAdd(&A.energy[kk], ke);
        if (kk >= 0 && (kfx != T(0) || kfy != T(0) || kfz != T(0) || kdc != T(0))) {
          atomicAdd(&A.force[3 * kk], kfx);
          atomicAdd(&A.force[3 * kk + 1], kfy);
          atomicAdd(&A.force[3 * kk + 2], kfz);
          atomicAdd(&A.dcn[kk], kdc);
        }
        }
      }
      __syncthreads();  // everybody is done with this buffer: it may be refilled
      tx = nx, ty = ny, nb ^= 1;
    }
  }
#undef TST
}

template <typename T>
constexpr size_t atm_grad_smem() {
  return sizeof(T) * (2 * GROUP * AVEC + 5 * GROUP + 2 * 2 * TILE * (AVEC + 1) + 5 * 2 * TILE +
                      2 * GROUP * GSTASH * TILE) +
         sizeof(int) * (2 * TILE + 2 * TILE + GROUP + 4);
}
template <typename T>
constexpr size_t twobody_grad_smem() {
  return sizeof(T) * (3 * NFREQ * 128 + 64 * AVEC + 5 * 64) + sizeof(int) * 64;
}

}  // namespace d4b200

using namespace d4b200;

namespace {

size_t al256(size_t x) { return (x + 255) / 256 * 256; }

struct LargeCarve {
  size_t cn, wq, w0, aq, a0, ulist, umask, ucount, cstash, queue, total;
  size_t zgd, z0gd, dzg, daq_cn, da0_cn, daq_q;
};

LargeCarve large_carve(int nat, size_t elem, int nctas, bool grad = false) {
  LargeCarve c;
  const size_t ng = (nat + GROUP - 1) / GROUP;
  size_t o = 0;
  c.queue = o, o += 256;  // status + queue
  c.cn = o, o += al256(nat * elem);
  c.wq = o, o += al256((size_t)nat * NREF * elem);
  c.w0 = o, o += al256((size_t)nat * NREF * elem);
  c.aq = o, o += al256((size_t)nat * AVEC * elem);
  c.a0 = o, o += al256((size_t)nat * AVEC * elem);
  c.ucount = o, o += al256(ng * sizeof(int));
  c.ulist = o, o += al256(ng * nat * sizeof(int));
  c.umask = o, o += al256(ng * nat * sizeof(unsigned));
  {
    size_t per_cta = (size_t)GROUP * (grad ? GSTASH : 3) * nat * elem;
    if (grad) {  // large_atm_grad_pipe: tiles of TILE list atoms with everything a tile pair needs
      const size_t nt = (nat + TILE - 1) / TILE;
      const size_t tile = elem == 8 ? atm_pipe_tile_bytes<double>() : atm_pipe_tile_bytes<float>();
      if (nt * tile > per_cta) per_cta = nt * tile;
    }
    c.cstash = o, o += al256((size_t)nctas * per_cta);
  }
  c.zgd = c.z0gd = c.dzg = c.daq_cn = c.da0_cn = c.daq_q = 0;
  if (grad) {
    c.zgd = o, o += al256((size_t)nat * NREF * elem);
    c.z0gd = o, o += al256((size_t)nat * NREF * elem);
    c.dzg = o, o += al256((size_t)nat * NREF * elem);
    c.daq_cn = o, o += al256((size_t)nat * AVEC * elem);
    c.da0_cn = o, o += al256((size_t)nat * AVEC * elem);
    c.daq_q = o, o += al256((size_t)nat * AVEC * elem);
  }
  c.total = o;
  return c;
}

template <typename T>
Par<T> large_par(const d4b200_params* p, double ga) {
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

int large_ctas(const d4b200_tables* h) { return h->num_sms * 2; }

template <typename T>
int run_large(d4b200_tables* h, const d4b200_params* par, int nat, const int64_t* numbers,
              const T* pos, const T* q, int row_begin, int row_end, int group_begin, int group_end,
              T* energy, T* cn_out, int* group_cost_out, void* ws, size_t ws_bytes,
              cudaStream_t st) {
  if (!h || !par || !numbers || !pos || !q || !ws || nat <= 0) return D4B200_EINVAL;
  if (par->model != D4B200_MODEL_D4) return D4B200_EPARAM;
  const int nctas = large_ctas(h);
  const LargeCarve c = large_carve(nat, sizeof(T), nctas);
  if (ws_bytes < c.total) return D4B200_EWORKSPACE;
  unsigned char* w = reinterpret_cast<unsigned char*>(ws);
  const int ng = (nat + GROUP - 1) / GROUP;
  LargeArgs<T> A;
  A.numbers = numbers;
  A.pos = pos;
  A.q = q;
  A.energy = energy;
  A.status = reinterpret_cast<int*>(w);
  A.queue = reinterpret_cast<int*>(w) + 1;
  A.cn = reinterpret_cast<T*>(w + c.cn);
  A.wq = reinterpret_cast<T*>(w + c.wq);
  A.w0 = reinterpret_cast<T*>(w + c.w0);
  A.aq = reinterpret_cast<T*>(w + c.aq);
  A.a0 = reinterpret_cast<T*>(w + c.a0);
  A.ucount = reinterpret_cast<int*>(w + c.ucount);
  A.ulist = reinterpret_cast<int*>(w + c.ulist);
  A.umask = reinterpret_cast<unsigned*>(w + c.umask);
  A.cstash = reinterpret_cast<T*>(w + c.cstash);
  A.gin = nullptr;
  A.zgd = A.z0gd = A.dzg = A.daq_cn = A.da0_cn = A.daq_q = A.force = A.dcn = A.dq = nullptr;
  A.nat = nat;
  A.ngroups = ng;
  A.ucap = nat;
  A.row_begin = row_begin < 0 ? 0 : row_begin;
  A.row_end = row_end > nat ? nat : row_end;
  A.group_begin = group_begin < 0 ? 0 : group_begin;
  A.group_end = group_end > ng ? ng : group_end;
  if constexpr (sizeof(T) == 8) {
    A.tab = h->t64;
  } else {
    A.tab = h->t32;
  }
  A.par = large_par<T>(par, h->ga);

  cudaMemsetAsync(w, 0, 256, st);
  cudaMemsetAsync(A.cn, 0, nat * sizeof(T), st);
  const int cchunks = nat > 4096 ? 8 : 1;
  large_cn<T><<<dim3((nat + 127) / 128, cchunks), 128, 0, st>>>(A, cchunks);
  if (cn_out) cudaMemcpyAsync(cn_out, A.cn, nat * sizeof(T), cudaMemcpyDeviceToDevice, st);
  large_weights<T, false><<<(nat * 8 + 255) / 256, 256, 0, st>>>(A);
  large_avec<T, false><<<(nat * AVEC + 255) / 256, 256, 0, st>>>(A);
  if (energy && A.row_end > A.row_begin) {
    const int rows = A.row_end - A.row_begin;
    large_twobody<T><<<dim3((rows + 127) / 128, cchunks), 128, 0, st>>>(A, cchunks);
  }
  if (A.par.has_atm && (group_cost_out || (energy && A.group_end > A.group_begin))) {
    large_union<T><<<ng, 256, 0, st>>>(A);
    if (group_cost_out)
      cudaMemcpyAsync(group_cost_out, A.ucount, ng * sizeof(int), cudaMemcpyDeviceToDevice, st);
    if (energy && A.group_end > A.group_begin) {
      // about six work items per resident CTA: groups are cut into row-tile slices when a
      // rank owns too few of them (multi-GPU runs, mid-size systems)
      const int ngl = A.group_end - A.group_begin;
      A.nslice = (6 * nctas + ngl - 1) / ngl;
      if (A.nslice > 16) A.nslice = 16;
      int grid = nctas;
      if (grid > ngl * A.nslice) grid = ngl * A.nslice;
      large_atm<T><<<grid, 256, 0, st>>>(A);
    }
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}


template <typename T>
void fill_common(LargeArgs<T>& A, d4b200_tables* h, const d4b200_params* par, const LargeCarve& c,
                 unsigned char* w, int nat, const int64_t* numbers, const T* pos, const T* q) {
  const int ng = (nat + GROUP - 1) / GROUP;
  A.numbers = numbers;
  A.pos = pos;
  A.q = q;
  A.energy = nullptr;
  A.status = reinterpret_cast<int*>(w);
  A.queue = reinterpret_cast<int*>(w) + 1;
  A.cn = reinterpret_cast<T*>(w + c.cn);
  A.wq = reinterpret_cast<T*>(w + c.wq);
  A.w0 = reinterpret_cast<T*>(w + c.w0);
  A.aq = reinterpret_cast<T*>(w + c.aq);
  A.a0 = reinterpret_cast<T*>(w + c.a0);
  A.ucount = reinterpret_cast<int*>(w + c.ucount);
  A.ulist = reinterpret_cast<int*>(w + c.ulist);
  A.umask = reinterpret_cast<unsigned*>(w + c.umask);
  A.cstash = reinterpret_cast<T*>(w + c.cstash);
  A.zgd = reinterpret_cast<T*>(w + c.zgd);
  A.z0gd = reinterpret_cast<T*>(w + c.z0gd);
  A.dzg = reinterpret_cast<T*>(w + c.dzg);
  A.daq_cn = reinterpret_cast<T*>(w + c.daq_cn);
  A.da0_cn = reinterpret_cast<T*>(w + c.da0_cn);
  A.daq_q = reinterpret_cast<T*>(w + c.daq_q);
  A.gin = nullptr;
  A.force = A.dcn = A.dq = nullptr;
  A.nat = nat;
  A.ngroups = ng;
  A.ucap = nat;
  if constexpr (sizeof(T) == 8) {
    A.tab = h->t64;
  } else {
    A.tab = h->t32;
  }
  A.par = large_par<T>(par, h->ga);
}

int large_grad_ctas(const d4b200_tables* h) { return h->num_sms; }

// Stage 1 of the gradient: everything except the CN chain rule.  Accumulates into
// force [nat,3], dcn [nat], dq [nat] (zero them first).
template <typename T>
int run_large_grad(d4b200_tables* h, const d4b200_params* par, int nat, const int64_t* numbers,
                   const T* pos, const T* q, const T* gin, int row_begin, int row_end,
                   int group_begin, int group_end, T* force, T* dcn, T* dq, void* ws,
                   size_t ws_bytes, cudaStream_t st, T* energy = nullptr) {
  if (!h || !par || !numbers || !pos || !q || !ws || !force || !dcn || !dq || nat <= 0) return D4B200_EINVAL;
  if (par->model != D4B200_MODEL_D4) return D4B200_EPARAM;
  const int nctas = large_grad_ctas(h);
  const LargeCarve c = large_carve(nat, sizeof(T), nctas, true);
  if (ws_bytes < c.total) return D4B200_EWORKSPACE;
  unsigned char* w = reinterpret_cast<unsigned char*>(ws);
  LargeArgs<T> A;
  fill_common<T>(A, h, par, c, w, nat, numbers, pos, q);
  const int ng = A.ngroups;
  A.gin = gin;
  A.energy = energy;  // optional: partial atomic energies of this rank's ranges from the same launches
  A.force = force;
  A.dcn = dcn;
  A.dq = dq;
  A.row_begin = row_begin < 0 ? 0 : row_begin;
  A.row_end = row_end > nat ? nat : row_end;
  A.group_begin = group_begin < 0 ? 0 : group_begin;
  A.group_end = group_end > ng ? ng : group_end;

  static bool configured[2] = {false, false};
  constexpr int dt = sizeof(T) == 8 ? 0 : 1;
  if (!configured[dt]) {
    cudaFuncSetAttribute(large_atm_grad<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)atm_grad_smem<T>());
    cudaFuncSetAttribute(large_atm_grad_pipe<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)atm_pipe_smem<T>());
    cudaFuncSetAttribute(large_twobody_grad<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)twobody_grad_smem<T>());
    configured[dt] = true;
  }
  cudaMemsetAsync(w, 0, 256, st);
  cudaMemsetAsync(A.cn, 0, nat * sizeof(T), st);
  const int cchunks = nat > 4096 ? 8 : 1;
  large_cn<T><<<dim3((nat + 127) / 128, cchunks), 128, 0, st>>>(A, cchunks);
  large_weights<T, true><<<(nat * 8 + 255) / 256, 256, 0, st>>>(A);
  large_avec<T, true><<<(nat * AVEC + 255) / 256, 256, 0, st>>>(A);
  if (A.row_end > A.row_begin) {
    const int rows = A.row_end - A.row_begin;
    large_twobody_grad<T><<<dim3((rows + 127) / 128, cchunks), 128, twobody_grad_smem<T>(), st>>>(A, cchunks);
  }
  if (A.par.has_atm && A.group_end > A.group_begin) {
    large_union<T><<<ng, 256, 0, st>>>(A);
    const int ngl = A.group_end - A.group_begin;
    A.nslice = (6 * nctas + ngl - 1) / ngl;
    if (A.nslice > 16) A.nslice = 16;
    int grid = nctas;
    if (grid > ngl * A.nslice) grid = ngl * A.nslice;
    static const bool pipe = [] {
      const char* v = getenv("D4B200_LARGE_PIPE");
      return v ? v[0] != '0' : D4_LARGE_PIPE_DEFAULT;
    }();
    if (pipe)
      large_atm_grad_pipe<T><<<grid, 512, atm_pipe_smem<T>(), st>>>(A);
    else
      large_atm_grad<T><<<grid, 512, atm_grad_smem<T>(), st>>>(A);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

// Stage 2: CN chain rule for the rows of this rank with the TOTAL dL/dcn.
template <typename T>
int run_large_chain(d4b200_tables* h, const d4b200_params* par, int nat, const int64_t* numbers,
                    const T* pos, const T* dcn_total, int row_begin, int row_end, T* force,
                    cudaStream_t st) {
  if (!h || !par || !numbers || !pos || !dcn_total || !force || nat <= 0) return D4B200_EINVAL;
  LargeArgs<T> A;
  LargeCarve c = large_carve(nat, sizeof(T), 1, true);
  unsigned char dummy = 0;
  fill_common<T>(A, h, par, c, &dummy, nat, numbers, pos, pos);
  A.dcn = const_cast<T*>(dcn_total);
  A.force = force;
  A.row_begin = row_begin < 0 ? 0 : row_begin;
  A.row_end = row_end > nat ? nat : row_end;
  if (A.row_end <= A.row_begin) return 0;
  const int cchunks = nat > 4096 ? 8 : 1;
  const int rows = A.row_end - A.row_begin;
  large_cn_chain<T><<<dim3((rows + 127) / 128, cchunks), 128, 0, st>>>(A, cchunks);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace

extern "C" {

int d4b200_large_group_size(void) { return GROUP; }

size_t d4b200_large_workspace_bytes(d4b200_tables_t h, int nat, int fp32) {
  if (!h || nat <= 0) return 0;
  const size_t e = large_carve(nat, fp32 ? 4 : 8, large_ctas(h)).total;
  const size_t g = large_carve(nat, fp32 ? 4 : 8, large_grad_ctas(h), true).total;
  return e > g ? e : g;
}

int d4b200_large_gradient_f64(d4b200_tables_t t, const d4b200_params* par, int nat,
                              const int64_t* numbers, const double* pos, const double* q,
                              const double* gin, int row_begin, int row_end, int group_begin,
                              int group_end, double* force, double* dcn, double* dq, void* ws,
                              size_t ws_bytes, void* stream) {
  return run_large_grad<double>(t, par, nat, numbers, pos, q, gin, row_begin, row_end, group_begin,
                                group_end, force, dcn, dq, ws, ws_bytes, (cudaStream_t)stream);
}
int d4b200_large_energy_gradient_f64(d4b200_tables_t t, const d4b200_params* par, int nat,
                                     const int64_t* numbers, const double* pos, const double* q,
                                     const double* gin, int row_begin, int row_end, int group_begin,
                                     int group_end, double* energy, double* force, double* dcn, double* dq,
                                     void* ws, size_t ws_bytes, void* stream) {
  return run_large_grad<double>(t, par, nat, numbers, pos, q, gin, row_begin, row_end, group_begin,
                                group_end, force, dcn, dq, ws, ws_bytes, (cudaStream_t)stream, energy);
}
int d4b200_large_energy_gradient_f32(d4b200_tables_t t, const d4b200_params* par, int nat,
                                     const int64_t* numbers, const float* pos, const float* q,
                                     const float* gin, int row_begin, int row_end, int group_begin,
                                     int group_end, float* energy, float* force, float* dcn, float* dq,
                                     void* ws, size_t ws_bytes, void* stream) {
  return run_large_grad<float>(t, par, nat, numbers, pos, q, gin, row_begin, row_end, group_begin,
                               group_end, force, dcn, dq, ws, ws_bytes, (cudaStream_t)stream, energy);
}
int d4b200_large_gradient_f32(d4b200_tables_t t, const d4b200_params* par, int nat,
                              const int64_t* numbers, const float* pos, const float* q,
                              const float* gin, int row_begin, int row_end, int group_begin,
                              int group_end, float* force, float* dcn, float* dq, void* ws,
                              size_t ws_bytes, void* stream) {
  return run_large_grad<float>(t, par, nat, numbers, pos, q, gin, row_begin, row_end, group_begin,
                               group_end, force, dcn, dq, ws, ws_bytes, (cudaStream_t)stream);
}
int d4b200_large_cn_chain_f64(d4b200_tables_t t, const d4b200_params* par, int nat,
                              const int64_t* numbers, const double* pos, const double* dcn_total,
                              int row_begin, int row_end, double* force, void* stream) {
  return run_large_chain<double>(t, par, nat, numbers, pos, dcn_total, row_begin, row_end, force,
                                 (cudaStream_t)stream);
}
int d4b200_large_cn_chain_f32(d4b200_tables_t t, const d4b200_params* par, int nat,
                              const int64_t* numbers, const float* pos, const float* dcn_total,
                              int row_begin, int row_end, float* force, void* stream) {
  return run_large_chain<float>(t, par, nat, numbers, pos, dcn_total, row_begin, row_end, force,
                                (cudaStream_t)stream);
}

int d4b200_large_energy_f64(d4b200_tables_t t, const d4b200_params* par, int nat,
                            const int64_t* numbers, const double* pos, const double* q,
                            int row_begin, int row_end, int group_begin, int group_end,
                            double* energy, double* cn, int* group_cost, void* ws, size_t ws_bytes,
                            void* stream) {
  return run_large<double>(t, par, nat, numbers, pos, q, row_begin, row_end, group_begin, group_end,
                           energy, cn, group_cost, ws, ws_bytes, (cudaStream_t)stream);
}
int d4b200_large_energy_f32(d4b200_tables_t t, const d4b200_params* par, int nat,
                            const int64_t* numbers, const float* pos, const float* q, int row_begin,
                            int row_end, int group_begin, int group_end, float* energy, float* cn,
                            int* group_cost, void* ws, size_t ws_bytes, void* stream) {
  return run_large<float>(t, par, nat, numbers, pos, q, row_begin, row_end, group_begin, group_end,
                          energy, cn, group_cost, ws, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"

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
#include "d4b200_small.cuh"  // math shims, d4_rcp, d4_zero_damp_arg, warp_sum

namespace d4b200 {

constexpr int GROUP = 16;   // centres per ATM group
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
  int nat, ngroups, ucap;
  int row_begin, row_end;      // two-body rows of this rank
  int group_begin, group_end;  // ATM centre groups of this rank
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
template <typename T>
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
  double S = 0.0;
  for (int k = 1; k <= rc; ++k) S += exp(-((double)k * arg - shift));
  double norm = S;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) norm += __shfl_xor_sync(0xffffffffu, norm, o);
  if (i < A.nat && a < NREF) {
    double gw = norm > 0.0 ? S / norm : 0.0, zeta = 0.0;
    if (rc > 0) {
      const double qmod = (double)A.q[i] + A.tab.zeff[z];
      if (qmod > 0.0) {
        const double scale = exp(A.tab.gamgc[z] * (1.0 - A.tab.refq[z * NREF + a] / (qmod - (double)d4_eps<T>())));
        zeta = exp(A.par.ga * (1.0 - scale));
      } else {
        zeta = exp(A.par.ga);
      }
    }
    A.wq[i * NREF + a] = (T)(zeta * gw);
    A.w0[i * NREF + a] = (T)((on ? A.tab.zeta0[z * NREF + a] : 0.0) * gw);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) large_avec(LargeArgs<T> A) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  const int i = t / AVEC, w = t - i * AVEC;
  if (i >= A.nat) return;
  T sq = T(0), s0 = T(0);
  const int zraw = (int)A.numbers[i];
  if (w < NFREQ && zraw > 0 && zraw < NELEM) {
    const T* al = A.tab.alpha_w + (size_t)zraw * NREF * NFREQ + w;
#pragma unroll
    for (int a = 0; a < NREF; ++a) {
      const T av = al[a * NFREQ];
      sq += A.wq[i * NREF + a] * av;
      s0 += A.w0[i * NREF + a] * av;
    }
  }
  A.aq[(size_t)i * AVEC + w] = sq;
  A.a0[(size_t)i * AVEC + w] = s0;
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
    if (tid == 0) gcur = A.group_begin + atomicAdd(A.queue, 1);
    __syncthreads();
    const int g = gcur;
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
        Pv = P.fac9 * d4_sqrt(fabs(c6)) * (rinv * rinv * rinv);
        uv = d4_zero_damp_arg(R0 * rinv, P.alp3, P.alp16 != 0);
      }
      cst[((size_t)j * 3 + 0) * A.ucap + u] = a;
      cst[((size_t)j * 3 + 1) * A.ucap + u] = Pv;
      cst[((size_t)j * 3 + 2) * A.ucap + u] = uv;
    }
    __syncthreads();

    const int ntile = (nU + TILE - 1) / TILE;
    for (int tx = 0; tx < ntile; ++tx) {
      for (int ty = tx; ty < ntile; ++ty) {
        __syncthreads();
        // ---- stage the two tiles (side 0 = X rows, side 1 = Y columns)
        for (int t = tid; t < 2 * TILE; t += 256) {
          const int side = t / TILE, l = t - side * TILE;
          const int u = (side == 0 ? tx : ty) * TILE + l;
          const bool ok = u < nU;
          const int x = ok ? list[u] : 0;
          tidx[side][l] = ok ? x : -1;
          tmask[side][l] = ok ? masks[u] : 0u;
          tpx[side][l] = ok ? A.pos[3 * x] : T(0);
          tpy[side][l] = ok ? A.pos[3 * x + 1] : T(0);
          tpz[side][l] = ok ? A.pos[3 * x + 2] : T(0);
          tsq[side][l] = ok ? A.tab.sqrt_r4r2[(int)A.numbers[x]] : T(0);
        }
        for (int t = tid; t < 2 * TILE * AVEC; t += 256) {
          const int side = t / (TILE * AVEC), r = t - side * TILE * AVEC;
          const int l = r / AVEC, w = r - l * AVEC;
          const int u = (side == 0 ? tx : ty) * TILE + l;
          tA0[side][l * (AVEC + 1) + w] = u < nU ? A.a0[(size_t)list[u] * AVEC + w] : T(0);
        }
        for (int t = tid; t < 2 * GROUP * 3 * TILE; t += 256) {
          const int side = t / (GROUP * 3 * TILE), r = t - side * GROUP * 3 * TILE;
          const int jc = r / TILE, l = r - jc * TILE;  // jc = j*3 + component
          const int u = (side == 0 ? tx : ty) * TILE + l;
          (&tst[side][0][0][0])[jc * TILE + l] = u < nU ? cst[(size_t)jc * A.ucap + u] : T(0);
        }
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
            const T Pik = P.fac9 * d4_sqrt(fabs(c6)) * (rinv * rinv * rinv);
            const T uik = d4_zero_damp_arg(R0 * rinv, P.alp3, P.alp16 != 0);
            unsigned both = mi & mk;
            T esum = T(0);
#pragma unroll 4
            for (int j = 0; j < GROUP; ++j) {
              if (both >> j & 1u) {
                const T a = tst[0][j][0][row], b = tst[1][j][0][lane];  // r_ji^2, r_jk^2
                const T X = a + b - c, Y = a - b + c, Z = b + c - a;
                const T abc = a * b * c;
                const T t = tst[0][j][2][row] * tst[1][j][2][lane] * uik;
                const T d = T(1) + T(6) * t;
                const T inv = d4_rcp(abc * d);
                esum += (T(0.375) * (X * Y * Z) * (inv * d) + T(1)) *
                        (tst[0][j][1][row] * tst[1][j][1][lane] * Pik * (inv * abc));
              }
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

}  // namespace d4b200

using namespace d4b200;

namespace {

size_t al256(size_t x) { return (x + 255) / 256 * 256; }

struct LargeCarve {
  size_t cn, wq, w0, aq, a0, ulist, umask, ucount, cstash, queue, total;
};

LargeCarve large_carve(int nat, size_t elem, int nctas) {
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
  c.cstash = o, o += al256((size_t)nctas * GROUP * 3 * nat * elem);
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
  large_weights<T><<<(nat * 8 + 255) / 256, 256, 0, st>>>(A);
  large_avec<T><<<(nat * AVEC + 255) / 256, 256, 0, st>>>(A);
  if (energy && A.row_end > A.row_begin) {
    const int rows = A.row_end - A.row_begin;
    large_twobody<T><<<dim3((rows + 127) / 128, cchunks), 128, 0, st>>>(A, cchunks);
  }
  if (A.par.has_atm && (group_cost_out || (energy && A.group_end > A.group_begin))) {
    large_union<T><<<ng, 256, 0, st>>>(A);
    if (group_cost_out)
      cudaMemcpyAsync(group_cost_out, A.ucount, ng * sizeof(int), cudaMemcpyDeviceToDevice, st);
    if (energy && A.group_end > A.group_begin) {
      int grid = nctas;
      if (grid > A.group_end - A.group_begin) grid = A.group_end - A.group_begin;
      large_atm<T><<<grid, 256, 0, st>>>(A);
    }
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace

extern "C" {

int d4b200_large_group_size(void) { return GROUP; }

size_t d4b200_large_workspace_bytes(d4b200_tables_t h, int nat, int fp32) {
  if (!h || nat <= 0) return 0;
  return large_carve(nat, fp32 ? 4 : 8, large_ctas(h)).total;
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

// tri_inv.cuh -- inverses of the IB x IB diagonal blocks of Teff (IB = 256 .. 2048) for the Float32 / Float16 solve leaves.
//
// The reference's solve leaf is a substitution with one barrier per pivot (src/trsm.jl:5-126) reached through 2n/256 - 1
// dependent launches (src/rectrxm.jl:101-198).  On B200 that dependency chain, not the flops, bounds the low-precision solves:
// with 128-wide leaves the K <= 512 updates and the leaves below them are ~6 % of the flops and ~45 % of the time.  So the
// diagonal blocks are inverted once per call up to order IB and every leaf becomes ONE triangular GEMM
//     X_blk = inv(Teff_blk) * V_blk
// (the standard GPU TRSM formulation; same flop count as the substitution it replaces).
//
// diag_prep_kernel (diag_prep.cuh) inverts the 128-blocks in FP64.  This file doubles them:
//     inv([T_XX 0; T_XY' ...]) : the off-diagonal block of the inverse of a pair (X = rows half, Y = columns half) is
//         Z = - inv(T_XX) * T(X,Y) * inv(T_YY)
// two batched GEMMs per level (phase 1: U = T(X,Y) * inv(Y), phase 2: Z = -inv(X) * U), every pair of the whole diagonal
// in one launch, triangular operands trimmed to their non-zero K range.  Arithmetic in Acc (FP32 for Float16 data, FP64 for
// Float32 data: the inverse is then correct to the element type's rounding, which is applied once by tri_inv_convert_kernel).
//
// Layout of the inverse workspace ("K-major", what the leaf GEMM's A operand (left side) / B operand (right side) wants):
//     W[g * IB + kl] = inv_i(g mod IB, kl),  g = global row index of Teff, i = g / IB, kl = column index local to block i.
#pragma once
#include "common.cuh"

namespace nla {

constexpr int TI_BM = 64, TI_BN = 64, TI_BK = 16, TI_THREADS = 256;

template <typename T> __device__ __forceinline__ double ti_load(const T* p);
template <> __device__ __forceinline__ double ti_load<double>(const double* p) { return *p; }
template <> __device__ __forceinline__ double ti_load<float>(const float* p) { return (double)*p; }
template <> __device__ __forceinline__ double ti_load<__half>(const __half* p) { return (double)__half2float(*p); }

template <typename T, typename Acc>
struct TriInvParams {
  const T* A; long long t_rs, t_cs;   // Teff(r,k) = A[r*t_rs + k*t_cs]
  int n;                              // order of Teff
  int ib;                             // order of the blocks being built (power of two >= 256)
  int s;                              // half size of this level: pairs of s-blocks are joined into 2s-blocks
  int lower;
  int pair0;                          // first pair handled by this launch (pair = blockIdx.y + pair0)
  Acc* W;                             // inverse workspace (pitch ib)
  Acc* U;                             // phase-1 products, same layout as W
};

// PHASE 1: U(X,Y) = T(X,Y) * inv(Y)       PHASE 2: W(X,Y) = -inv(X) * U(X,Y)
// grid.x = 64 x 64 tiles of one s x s block, grid.y = pair index along the whole diagonal.
template <typename T, typename Acc, int PHASE>
__global__ void __launch_bounds__(TI_THREADS) tri_inv_step_kernel(const TriInvParams<T, Acc> p) {
  __shared__ Acc As[TI_BK][TI_BM + 4];
  __shared__ Acc Bs[TI_BK][TI_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;   // thread computes rows tx*4.., cols ty*4..
  const int s = p.s;
  const int g0 = (blockIdx.y + p.pair0) * 2 * s;     // first global row of the pair
  const int blk_end = min(p.n, (g0 / p.ib + 1) * p.ib);
  const int h0 = g0, h1 = g0 + s;                    // first / second half
  const int n1 = min(s, blk_end - h1);               // extent of the second half (first half is full whenever n1 > 0)
  if (n1 <= 0) return;
  const int gx = p.lower ? h1 : h0, gy = p.lower ? h0 : h1;   // rows / columns of the off-diagonal block of the inverse
  const int nx = p.lower ? n1 : s, ny = p.lower ? s : n1;
  const int lx = gx % p.ib, ly = gy % p.ib;
  const int tiles_n = (s + TI_BN - 1) / TI_BN;
  const int m0 = (blockIdx.x / tiles_n) * TI_BM, n0 = (blockIdx.x % tiles_n) * TI_BN;
  if (m0 >= nx || n0 >= ny) return;
  const int K = PHASE == 1 ? ny : nx;
  // non-zero K range of the triangular operand
  int klo = 0, khi = K;
  if (PHASE == 1) {   // inv(Y)(k,j): lower -> k >= j, upper -> k <= j
    if (p.lower) klo = n0; else khi = min(K, n0 + TI_BN);
  } else {            // inv(X)(i,k): lower -> k <= i, upper -> k >= i
    if (p.lower) khi = min(K, m0 + TI_BM); else klo = m0;
  }
  klo = klo / TI_BK * TI_BK;

  Acc acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = Acc(0);

  const bool a_kcontig = PHASE == 2 || p.t_cs == 1;   // A(i,k): k contiguous?
  for (int k0 = klo; k0 < khi; k0 += TI_BK) {
#pragma unroll
    for (int l = 0; l < 4; l++) {
      const int idx = tid + l * TI_THREADS;
      int mi, ki;
      if (a_kcontig) { ki = idx & 15; mi = idx >> 4; } else { mi = idx & 63; ki = idx >> 6; }
      const int i = m0 + mi, k = k0 + ki;
      Acc av = Acc(0);
      if (i < nx && k < khi) {
        if (PHASE == 1) av = (Acc)ti_load<T>(p.A + (long long)(gx + i) * p.t_rs + (long long)(gy + k) * p.t_cs);
        else av = p.W[(long long)(gx + i) * p.ib + lx + k];
      }
      As[ki][mi] = av;
      // B(k,j): j contiguous in both phases
      const int nj = idx & 63, kj = idx >> 6;
      const int j = n0 + nj, k2 = k0 + kj;
      Acc bv = Acc(0);
      if (j < ny && k2 < khi) {
        if (PHASE == 1) bv = p.W[(long long)(gy + k2) * p.ib + ly + j];
        else bv = p.U[(long long)(gx + k2) * p.ib + ly + j];
      }
      Bs[kj][nj] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TI_BK; k++) {
      Acc a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[k][tx * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[k][ty * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
  Acc* out = PHASE == 1 ? p.U : p.W;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int row = m0 + tx * 4 + i;
    if (row >= nx) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int col = n0 + ty * 4 + j;
      if (col >= ny) continue;
      out[(long long)(gx + row) * p.ib + ly + col] = PHASE == 1 ? acc[i][j] : -acc[i][j];
    }
  }
}

// Rounds the finished inverses once to the element type: Wt[g*ib + kl] = (T) W[g*ib + kl] inside the triangle (at 128-block
// granularity; the diagonal 128-blocks carry explicit zeros in their other half), zero elsewhere and in the rows beyond n.
template <typename T, typename Acc>
__global__ void __launch_bounds__(256) tri_inv_convert_kernel(const Acc* __restrict__ W, T* __restrict__ Wt, int n, int row0, int row1, int ib, int lower) {
  const long long total = (long long)row1 * ib;   // rows [row0, row1) of the workspace
  for (long long e = (long long)row0 * ib + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(e / ib), kl = (int)(e % ib);
    const int rb = (g % ib) >> 7, kb = kl >> 7;
    const int bsz = min(ib, n - g / ib * ib);   // order of this (possibly ragged last) block: columns beyond it were never written
    const bool in = g < n && kl < bsz && (lower ? kb <= rb : kb >= rb);
    Traits<T>::st(Wt + e, in ? (typename Traits<T>::Acc)W[e] : (typename Traits<T>::Acc)0);
  }
}

}  // namespace nla

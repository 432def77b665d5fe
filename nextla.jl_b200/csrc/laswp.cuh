// laswp.cuh -- row interchanges on a device matrix: the device-side counterpart of the reference's `laswp`
// (src/lu.jl:470-530), which its recursive LU applies between the panel factorisation and the TRSM + GEMM update
// (src/lu.jl:274, :297).  With nla_trxm (unit-lower solve) and nla_gemm_update this puts every O(n^3) step of getrf2! on the
// device; since round 2 the panel factorisation is on the device too (getrf.cuh, nla_getrf2).
// laswp_kernel (backward walks): one thread per column walks the pivots in order (interchanges of one column are sequentially dependent, columns are
// independent); the pivot vector is read through the read-only path and broadcast to the warp.
#pragma once
#include "common.cuh"

namespace nla {

template <typename T>
__global__ void __launch_bounds__(256) laswp_kernel(T* __restrict__ A, long long lda, long long ncols, long long k1, long long k2,
                                                    const long long* __restrict__ ipiv, int reverse) {
  const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  T* c = A + col * lda;
  // rows and pivots are 1-based (Julia / LAPACK); ipiv[i-1] is the row exchanged with row i
  if (!reverse) {
    for (long long i = k1; i <= k2; i++) {
      const long long ip = __ldg(ipiv + (i - 1));
      if (ip != i) { const T t = c[i - 1]; c[i - 1] = c[ip - 1]; c[ip - 1] = t; }
    }
  } else {
    for (long long i = k2; i >= k1; i--) {
      const long long ip = __ldg(ipiv + (i - 1));
      if (ip != i) { const T t = c[i - 1]; c[i - 1] = c[ip - 1]; c[ip - 1] = t; }
    }
  }
}

// Forward interchanges, planned.  The walk above is latency-bound: every step's loads follow the previous step's stores, so a column
// advances one pivot per L2 round trip (ncu, recursive LU n = 16384: 124 ms in laswp, a third of the factorisation).  The interchanges
// are the same for every column, so their bookkeeping is done once per call instead of once per element:
//   * laswp_plan_kernel takes the steps in batches of LASWP_NB (one thread per batch, all batches in parallel).  The net effect of a
//     batch is "position d receives the value position s held before the batch" for at most 2 * LASWP_NB positions: the thread replays
//     its batch on POSITIONS (unrolled in registers: a later step sees what an earlier step of the same batch left at a position,
//     whichever of the two roles -- row or partner -- the position played), drops every store a later step of the batch overwrites,
//     and leaves per step: the partner row, the source of the row store and the source of the partner store (-1 = no store).
//   * laswp_apply_kernel moves the data: half a warp per column, lane j of it owns step j of the current batch -- two gathers, a warp
//     barrier, two scatters to distinct addresses.  The 16 rows of a batch are contiguous in a column, so that side of every access
//     is one 128-byte line per column instead of one line per element (the per-thread-per-column layout of the first planned version
//     was bound by the SM's load/store unit: 32 lines per instruction, 6 us per batch, 45 ms per factorisation).
constexpr int LASWP_NB = 16, LASWP_THREADS = 128;

__global__ void __launch_bounds__(128) laswp_plan_kernel(const long long* __restrict__ ipiv, long long k1, long long k2, int* __restrict__ plan_t,
                                                         int* __restrict__ plan_row, int* __restrict__ plan_tgt) {
  const long long steps = k2 - k1 + 1;
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b * LASWP_NB >= steps) return;
  const int r0 = (int)(k1 - 1 + b * LASWP_NB);              // 0-based row of the batch's first step
  const int nb = (int)min((long long)LASWP_NB, steps - b * LASWP_NB);
  int t[LASWP_NB], na[LASWP_NB], nt[LASWP_NB];
#pragma unroll
  for (int j = 0; j < LASWP_NB; j++) t[j] = j < nb ? (int)(ipiv[r0 + j] - 1) : r0 + j;     // beyond the last step: self (never stored)
#pragma unroll
  for (int j = 0; j < LASWP_NB; j++) {
    int ca = r0 + j, cb = t[j];                   // what the two positions hold: initially their own values
#pragma unroll
    for (int q = 0; q < j; q++) {                 // step q wrote row r0 + q <- na[q], then row t[q] <- nt[q]
      if (t[q] == r0 + j) ca = nt[q];
      if (r0 + q == t[j]) cb = na[q];
      if (t[q] == t[j]) cb = nt[q];
    }
    if (t[j] == r0 + j) cb = ca;
    na[j] = cb; nt[j] = ca;
  }
#pragma unroll
  for (int j = 0; j < LASWP_NB; j++) {
    bool row_live = j < nb, tgt_live = j < nb && t[j] != r0 + j;    // a self-interchange keeps its row store only
#pragma unroll
    for (int q = j + 1; q < LASWP_NB; q++) {
      if (q < nb) {
        if (t[q] == r0 + j) row_live = false;                       // a later partner store lands on this row
        if (t[q] == t[j] || r0 + q == t[j]) tgt_live = false;       // a later store (partner or row) lands on this partner row
      }
    }
    if (row_live && na[j] == r0 + j) row_live = false;              // value already in place
    if (tgt_live && nt[j] == t[j]) tgt_live = false;
    const long long e = b * LASWP_NB + j;
    plan_t[e] = t[j];
    plan_row[e] = row_live ? na[j] : -1;
    plan_tgt[e] = tgt_live ? nt[j] : -1;
  }
}

template <typename T>
__global__ void __launch_bounds__(LASWP_THREADS) laswp_apply_kernel(T* __restrict__ A, long long lda, long long ncols, long long k1, long long k2,
                                                                    const int* __restrict__ plan_t, const int* __restrict__ plan_row,
                                                                    const int* __restrict__ plan_tgt) {
  const int lane = threadIdx.x & 31, j = lane & (LASWP_NB - 1);
  const long long col = ((long long)blockIdx.x * (LASWP_THREADS / 32) + (threadIdx.x >> 5)) * 2 + (lane >> 4);
  const bool active = col < ncols;
  T* c = A + (active ? col : 0) * lda;
  const long long steps = k2 - k1 + 1;
  const long long nbatches = (steps + LASWP_NB - 1) / LASWP_NB;
  int t = plan_t[j], sr = plan_row[j], st = plan_tgt[j];
  for (long long b = 0; b < nbatches; b++) {
    const int r = (int)(k1 - 1 + b * LASWP_NB) + j;
    T vr = T(), vt = T();
    const bool lr = active && sr >= 0, lt = active && st >= 0;
    if (lr) vr = c[sr];
    if (lt) vt = c[st];
    const int t_now = t;
    if (b + 1 < nbatches) {                       // next batch's plan while the gathers are in flight
      const long long e = (b + 1) * LASWP_NB + j;
      t = plan_t[e]; sr = plan_row[e]; st = plan_tgt[e];
    }
    __syncwarp();                                 // every gather of the batch before any scatter
    if (lr) c[r] = vr;
    if (lt) c[t_now] = vt;
    __syncwarp();                                 // the batch's stores before the next batch's gathers
  }
}

}  // namespace nla

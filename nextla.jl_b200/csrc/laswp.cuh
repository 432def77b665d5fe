// laswp.cuh -- row interchanges on a device matrix: the device-side counterpart of the reference's `laswp`
// (src/lu.jl:470-530), which its recursive LU applies between the panel factorisation and the TRSM + GEMM update
// (src/lu.jl:274, :297).  With nla_trxm (unit-lower solve) and nla_gemm_update this puts every O(n^3) step of getrf2! on the
// device; the panel factorisation itself stays with the caller.
// One thread per column walks the pivots in order (interchanges of one column are sequentially dependent, columns are
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

}  // namespace nla

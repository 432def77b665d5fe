// gemm_simt.cuh -- generic strided GEMM update  C <- post * (beta*C + sgn * A*B)  for any element type,
// any (row stride, column stride) operand views and any alignment.  Used (a) for FP64 problems the TMA
// path cannot take (odd leading dimension, sizes not a multiple of 8, unaligned base) and (b) as the
// round-1 path for Float32 / Float16 until their tcgen05 kernels land (see DESIGN.md "what comes next").
// Replaces src/matmul.jl:5-66 semantics (accumulate, then one rounded update of the output) with FP32
// accumulation for Float16 instead of the reference's FP16 accumulation.
#pragma once
#include "common.cuh"

namespace nla {

template <typename T>
struct GemmSimtParams {
  int M, N, K;
  const T* A; long long a_rs, a_cs;   // A(i,k) = A[i*a_rs + k*a_cs]
  const T* B; long long b_rs, b_cs;   // B(k,j) = B[k*b_rs + j*b_cs]
  T* C; long long ldc;                // column-major output
  double beta, sgn, post;
  int tri_mode;    // 0: whole block; 1 / 2: only the lower / upper triangle (diagonal included) of C is READ AND WRITTEN (nla_lauum)
  int overwrite;   // 1: C <- post * sgn * A*B (old contents not read)
};

constexpr int GS_BM = 64, GS_BN = 64, GS_BK = 16, GS_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(GS_THREADS) gemm_simt_kernel(const GemmSimtParams<T> p) {
  using Acc = typename Traits<T>::Acc;
  __shared__ Acc As[GS_BK][GS_BM + 4];
  __shared__ Acc Bs[GS_BK][GS_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // thread computes rows tx*4.., cols ty*4..
  const int m0 = blockIdx.x * GS_BM, n0 = blockIdx.y * GS_BN;
  Acc acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = Acc(0);

  const bool a_mcontig = (p.a_rs == 1), b_kcontig = (p.b_rs == 1);
  for (int k0 = 0; k0 < p.K; k0 += GS_BK) {
#pragma unroll
    for (int l = 0; l < 4; l++) {
      const int idx = tid + l * GS_THREADS;
      int mi, ki;
      if (a_mcontig) { mi = idx & 63; ki = idx >> 6; } else { ki = idx & 15; mi = idx >> 4; }
      const int gm = m0 + mi, gk = k0 + ki;
      As[ki][mi] = (gm < p.M && gk < p.K) ? Traits<T>::ld(p.A + (long long)gm * p.a_rs + (long long)gk * p.a_cs) : Acc(0);
      int ni, kj;
      if (b_kcontig) { kj = idx & 15; ni = idx >> 4; } else { ni = idx & 63; kj = idx >> 6; }
      const int gn = n0 + ni, gk2 = k0 + kj;
      Bs[kj][ni] = (gn < p.N && gk2 < p.K) ? Traits<T>::ld(p.B + (long long)gk2 * p.b_rs + (long long)gn * p.b_cs) : Acc(0);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GS_BK; k++) {
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
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int col = n0 + ty * 4 + j;
    if (col >= p.N) continue;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int row = m0 + tx * 4 + i;
      if (row >= p.M) continue;
      if (p.tri_mode == 1 ? row < col : (p.tri_mode == 2 && row > col)) continue;
      T* cp = p.C + (long long)col * p.ldc + row;
      if (p.overwrite) { Traits<T>::st(cp, (Acc)(p.post * p.sgn * (double)acc[i][j])); continue; }
      // mirror the reference's rounding points: scale rounded to T, update evaluated in Float64 and
      // rounded once (src/matmul.jl:64), final scale rounded to T (src/rectrxm.jl:64,72)
      double v = (double)Traits<T>::ld(cp);
      if (p.beta != 1.0) { Traits<T>::st(cp, (Acc)(p.beta * v)); v = (double)Traits<T>::ld(cp); }
      v = v + p.sgn * (double)acc[i][j];
      if (p.post != 1.0) { Traits<T>::st(cp, (Acc)v); v = p.post * (double)Traits<T>::ld(cp); }
      Traits<T>::st(cp, (Acc)v);
    }
  }
}

}  // namespace nla

// diag_prep.cuh -- diagonal-block preparation for the Float32 / Float16 tensor-core leaves.
//
// The reference's leaves solve with one barrier per pivot (src/trsm.jl:5-126) or multiply 16x16 tiles (src/trmm.jl:43-312).
// On B200 a 128 x 128 triangular block applied to thousands of right-hand sides is a GEMM-shaped job, so the low-precision
// leaves run on the tensor cores too:
//   solve:     X_blk = inv(Teff_blk) * (alpha * B_blk)     multiply:  Y_blk = tri(Teff_blk) * B_blk
// This kernel builds, once per call and for every 128-block of the diagonal, the operand P those leaf GEMMs consume:
//   P = inv(Teff_blk) (computed in FP64 from the stored values, rounded once to T) or P = Teff_blk masked to its triangle
// (the opposite triangle of A is never read and may hold anything).  P is written K-major -- W[(128*i + r)*128 + k] =
// P_i(r, k) -- so that it is the A operand of the left-side leaf and the B operand of the right-side leaf (gemm_tc.cuh).
// Rows/columns beyond a ragged last block are zero.
//
// One CTA (128 threads) per block.  Upper blocks are reversed to lower ones on load and reversed back on store.  Thread j
// owns column j of the inverse: forward substitution in FP64 with the triangular block (exact in FP32) and the growing
// inverse in shared memory; all threads read the same L(r,k) (broadcast) and their own column (conflict free).
#pragma once
#include "common.cuh"

namespace nla {

constexpr int DP_B = 128;                       // diagonal block order == GEMM M tile == recursion cutoff of the TC path
constexpr int DP_LP = DP_B + 1;
constexpr int DP_SMEM_BYTES = DP_B * DP_LP * 4 + DP_B * DP_LP * 8;

template <typename T>
struct DiagPrepParams {
  const T* A; long long t_rs, t_cs;   // Teff(r,k) = A[r*t_rs + k*t_cs]
  int n;                              // order of Teff
  int lower, solve;
  int block0;                         // first diagonal block handled by this launch (block b = blockIdx.x + block0)
  T* W;                               // K-major workspace, 128 x 128 per block
};

template <typename T>
__global__ void __launch_bounds__(DP_B) diag_prep_kernel(const DiagPrepParams<T> p) {
  extern __shared__ __align__(16) uint8_t dp_smem[];
  float* Ls = reinterpret_cast<float*>(dp_smem);                       // [r][k], pitch DP_LP (normalised to lower)
  double* Xs = reinterpret_cast<double*>(dp_smem + DP_B * DP_LP * 4);  // [r][j], pitch DP_LP
  const int tid = threadIdx.x;
  const int b = blockIdx.x + p.block0;
  const int off = b * DP_B;
  const int t = min(DP_B, p.n - off);
  const T* Ab = p.A + (long long)off * (p.t_rs + p.t_cs);

  // ---- load the block, masked to its triangle, normalised to lower by index reversal ----
  const bool r_contig = (p.t_rs == 1);
  for (int o = 0; o < DP_B; o++) {
    const int r = r_contig ? tid : o, k = r_contig ? o : tid;   // thread index along the contiguous direction of A
    float v = 0.f;
    if (r < t && k < t) {
      const int R = p.lower ? r : t - 1 - r, K = p.lower ? k : t - 1 - k;
      if (k <= r) v = (float)Traits<T>::ld(Ab + (long long)R * p.t_rs + (long long)K * p.t_cs);
    } else if (r == k) {
      v = 1.f;   // identity padding keeps the substitution finite; padded entries are written as zeros below
    }
    Ls[r * DP_LP + k] = v;
  }
  __syncthreads();

  if (p.solve) {
    // ---- column `tid` of inv(L): x_r = (delta_{r,tid} - sum_{k=tid}^{r-1} L(r,k) x_k) / L(r,r) ----
    const int j = tid;
    for (int r = 0; r < j; r++) Xs[r * DP_LP + j] = 0.0;
    for (int r = j; r < DP_B; r++) {
      double s0 = (r == j) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      const float* Lr = Ls + r * DP_LP;
      int k = j;
      for (; k + 3 < r; k += 4) {
        s0 -= (double)Lr[k] * Xs[k * DP_LP + j];
        s1 -= (double)Lr[k + 1] * Xs[(k + 1) * DP_LP + j];
        s2 -= (double)Lr[k + 2] * Xs[(k + 2) * DP_LP + j];
        s3 -= (double)Lr[k + 3] * Xs[(k + 3) * DP_LP + j];
      }
      for (; k < r; k++) s0 -= (double)Lr[k] * Xs[k * DP_LP + j];
      Xs[r * DP_LP + j] = ((s0 + s1) + (s2 + s3)) / (double)Lr[r];
    }
  } else {
    for (int r = 0; r < DP_B; r++) Xs[r * DP_LP + tid] = (double)Ls[r * DP_LP + tid];
  }
  __syncthreads();

  // ---- store K-major (k contiguous), reversing back for upper blocks ----
  T* Wb = p.W + (long long)off * DP_B;
  for (int r = 0; r < DP_B; r++) {
    const int k = tid;
    float v = 0.f;
    if (r < t && k < t) {
      const int rr = p.lower ? r : t - 1 - r, kk = p.lower ? k : t - 1 - k;   // position in the normalised (lower) block
      if (kk <= rr) v = (float)Xs[rr * DP_LP + kk];
    }
    Traits<T>::st(Wb + (long long)r * DP_B + k, v);
  }
}

}  // namespace nla

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
// One CTA (1024 threads) per block.  Upper blocks are reversed to lower ones on load and reversed back on store.  Eight
// lanes share column j of the inverse: forward substitution in FP64 with the triangular block (exact in FP32) and the growing
// inverse in shared memory; the dot product of every row is split eight ways and reduced with warp shuffles, which cuts the
// serial chain from ~8000 dependent FMAs (one thread per column: 250 us measured) to ~1000.
#pragma once
#include "common.cuh"

namespace nla {

constexpr int DP_B = 128;                       // diagonal block order == GEMM M tile == recursion cutoff of the TC path
constexpr int DP_LP = DP_B + 1;
constexpr int DP_THREADS = 1024, DP_SPLIT = DP_THREADS / DP_B;   // lanes per column
constexpr int DP_SMEM_BYTES = DP_B * DP_LP * 4 + DP_B * DP_LP * 8;

template <typename TO> __device__ __forceinline__ void dp_store(TO* p, double v);
template <> __device__ __forceinline__ void dp_store<double>(double* p, double v) { *p = v; }
template <> __device__ __forceinline__ void dp_store<float>(float* p, double v) { *p = (float)v; }
template <> __device__ __forceinline__ void dp_store<__half>(__half* p, double v) { *p = __float2half_rn((float)v); }

// TO = element type of the workspace: T itself (128-wide leaves consume W directly), or the accumulation type of the block-inverse
// doubling (tri_inv.cuh), which places block b at rows [128b, 128b+128), local columns [(128b) mod ib, +128) of a pitch-`ib` workspace.
template <typename T, typename TO = T>
struct DiagPrepParams {
  const T* A; long long t_rs, t_cs;   // Teff(r,k) = A[r*t_rs + k*t_cs]
  int n;                              // order of Teff
  int lower, solve;
  int unit;                           // 1: unit diagonal (stored diagonal not read)
  int block0;                         // first diagonal block handled by this launch (block b = blockIdx.x + block0)
  TO* W;                              // K-major workspace, 128 x 128 per block
  int pitch, ib;                      // elements between consecutive rows of W; order of the enclosing inverse block (128: none)
};

template <typename T, typename TO = T>
__global__ void __launch_bounds__(DP_THREADS) diag_prep_kernel(const DiagPrepParams<T, TO> p) {
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
  for (int e = tid; e < DP_B * DP_B; e += DP_THREADS) {
    const int fast = e % DP_B, slow = e / DP_B;                   // `fast` runs along the contiguous direction of A
    const int r = r_contig ? fast : slow, k = r_contig ? slow : fast;
    float v = 0.f;
    if (r < t && k < t) {
      const int R = p.lower ? r : t - 1 - r, K = p.lower ? k : t - 1 - k;
      if (k <= r) v = (p.unit && k == r) ? 1.f : (float)Traits<T>::ld(Ab + (long long)R * p.t_rs + (long long)K * p.t_cs);
    } else if (r == k) {
      v = 1.f;   // identity padding keeps the substitution finite; padded entries are written as zeros below
    }
    Ls[r * DP_LP + k] = v;
  }
  __syncthreads();

  if (p.solve) {
    // ---- column j of inv(L): x_r = (delta_{r,j} - sum_{k=j}^{r-1} L(r,k) x_k) / L(r,r); lanes part = 0..7 split the sum ----
    const int j = tid / DP_SPLIT, part = tid % DP_SPLIT;
    for (int r = part; r < j; r += DP_SPLIT) Xs[r * DP_LP + j] = 0.0;
    const int j0 = (tid / 32) * (32 / DP_SPLIT);   // first column of this warp: every lane runs the same trip count (full-mask shuffles)
    for (int r = j0; r < DP_B; r++) {
      const float* Lr = Ls + r * DP_LP;
      double s = 0.0;
      for (int k = j + part; k < r; k += DP_SPLIT) s -= (double)Lr[k] * Xs[k * DP_LP + j];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (part == 0 && r >= j) Xs[r * DP_LP + j] = (s + ((r == j) ? 1.0 : 0.0)) / (double)Lr[r];
      __syncwarp();
    }
  } else {
    for (int e = tid; e < DP_B * DP_B; e += DP_THREADS) Xs[(e / DP_B) * DP_LP + e % DP_B] = (double)Ls[(e / DP_B) * DP_LP + e % DP_B];
  }
  __syncthreads();

  // ---- store K-major (k contiguous), reversing back for upper blocks ----
  TO* Wb = p.W + (long long)off * p.pitch + (off % p.ib);
  for (int e = tid; e < DP_B * DP_B; e += DP_THREADS) {
    const int k = e % DP_B, r = e / DP_B;
    double vd = 0.0;
    if (r < t && k < t) {
      const int rr = p.lower ? r : t - 1 - r, kk = p.lower ? k : t - 1 - k;   // position in the normalised (lower) block
      if (kk <= rr) vd = Xs[rr * DP_LP + kk];
    }
    dp_store<TO>(Wb + (long long)r * p.pitch + k, vd);
  }
}

}  // namespace nla

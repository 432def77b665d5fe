// complex.cuh -- ComplexF32 / ComplexF64 support for the recursive TRSM / TRMM path.
//
// The reference advertises complex element types (README.md:20) and builds `Adjoint(A)` for transpose = 'C' (src/rectrxm.jl:57), but
// `unified_rec` is restricted to `T <: AbstractFloat` (:101), so a complex call fails there.  Here it works, with 'C' != 'T':
//   * the interleaved (re, im) matrices of the caller are split into PLANAR real matrices in a library workspace (A once per call,
//     conjugated on the fly for 'C'; B with alpha folded in for a solve -- the reference's own `B .= alpha .* B` pass, :64),
//   * the reference's recursion (same schedule as the real path, cutoff 128) runs on the planes: a complex update
//         C -+= T * X   is   Cr -+= Tr Xr - Ti Xi ,  Ci -+= Tr Xi + Ti Xr
//     i.e. FOUR real GEMM updates on the tensor cores (FP64 DMMA / 3xTF32 tcgen05) through the same kernels as the real path,
//   * a diagonal block is solved / multiplied by ctri_kernel below: complex substitution (the reference's scaled form, src/trsm.jl:15-27,
//     in complex arithmetic) on the planes, one thread per right-hand-side vector, the block's rows staged in chunks of 16,
//   * the planes of B are merged back into the caller's interleaved matrix (alpha folded in for a multiply, :72).
#pragma once
#include "common.cuh"

namespace nla {

// Z (interleaved, rows x cols, leading dimension ldz complex elements) -> planes Pr, Pi (leading dimension ldp);
// P = (sr + i si) * Z, imaginary part negated afterwards when conj != 0.
template <typename R>
__global__ void __launch_bounds__(256) cplx_split_kernel(const R* __restrict__ Z, long long ldz, R* __restrict__ Pr, R* __restrict__ Pi, long long ldp,
                                                         long long rows, long long cols, double sr, double si, int conj) {
  const long long total = rows * cols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e % rows, c = e / rows;
    const R zr = Z[2 * (r + c * ldz)], zi = Z[2 * (r + c * ldz) + 1];
    R pr = zr, pi = zi;
    if (sr != 1.0 || si != 0.0) { pr = (R)(sr * (double)zr - si * (double)zi); pi = (R)(sr * (double)zi + si * (double)zr); }
    Pr[r + c * ldp] = pr;
    Pi[r + c * ldp] = conj ? -pi : pi;
  }
}

template <typename R>
__global__ void __launch_bounds__(256) cplx_merge_kernel(R* __restrict__ Z, long long ldz, const R* __restrict__ Pr, const R* __restrict__ Pi, long long ldp,
                                                         long long rows, long long cols, double sr, double si) {
  const long long total = rows * cols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e % rows, c = e / rows;
    R pr = Pr[r + c * ldp], pi = Pi[r + c * ldp];
    if (sr != 1.0 || si != 0.0) { const R tr = (R)(sr * (double)pr - si * (double)pi); pi = (R)(sr * (double)pi + si * (double)pr); pr = tr; }
    Z[2 * (r + c * ldz)] = pr;
    Z[2 * (r + c * ldz) + 1] = pi;
  }
}

constexpr int CT_THREADS = 128, CT_RB = 16, CT_KC = 32;

template <typename R>
struct CTriParams {
  const R* Ar; const R* Ai; long long t_rs, t_cs;   // planes of Teff: Teff(r,k) = A[r*t_rs + k*t_cs] (block origin included by the caller)
  int sz, unit;
  R* Vr; R* Vi; long long es, vs;                   // element i of vector v at V[i*es + v*vs] (block origin included by the caller)
  int nv;
};

// Diagonal-block leaf on planar complex data: SOLVE: x <- Teff^-1 x;  otherwise x <- Teff x.  One thread per vector, 16 rows at a time.
template <typename R, bool LOWER, bool SOLVE>
__global__ void __launch_bounds__(CT_THREADS) ctri_kernel(const CTriParams<R> p) {
  __shared__ R Tr[CT_KC][CT_RB], Ti[CT_KC][CT_RB];   // T[j][i] = Teff(r0 + i, k0 + j)
  const int tid = threadIdx.x;
  const int v = blockIdx.x * CT_THREADS + tid;
  const bool active = v < p.nv;
  R* Vr = p.Vr + (long long)v * p.vs;
  R* Vi = p.Vi + (long long)v * p.vs;
  const int nchunks = (p.sz + CT_RB - 1) / CT_RB;
  // a solve walks the unknowns forward (lower) / backward (upper); an in-place multiply walks the other way so that the entries it
  // still needs are the ones it has not overwritten yet
  constexpr bool ASC = (SOLVE == LOWER);
  for (int c = 0; c < nchunks; c++) {
    const int r0 = (ASC ? c : nchunks - 1 - c) * CT_RB;
    const int nr = min(CT_RB, p.sz - r0);
    R ar[CT_RB], ai[CT_RB];     // accumulators: solve: rhs - sum; multiply: sum
    R xr[CT_RB], xi[CT_RB];     // multiply: the chunk's own (old) entries
#pragma unroll
    for (int i = 0; i < CT_RB; i++) {
      R br = 0, bi = 0;
      if (active && i < nr) { br = Vr[(long long)(r0 + i) * p.es]; bi = Vi[(long long)(r0 + i) * p.es]; }
      xr[i] = br; xi[i] = bi;
      ar[i] = SOLVE ? br : R(0); ai[i] = SOLVE ? bi : R(0);
    }
    // off-chunk columns: k < r0 (lower) or k >= r0 + nr (upper) -- already solved (solve) / still original (multiply)
    const int kbeg = LOWER ? 0 : r0 + nr, kend = LOWER ? r0 : p.sz;
    for (int k0 = kbeg; k0 < kend; k0 += CT_KC) {
      const int nk = min(CT_KC, kend - k0);
      __syncthreads();
      for (int e = tid; e < CT_KC * CT_RB; e += CT_THREADS) {
        const int i = p.t_rs == 1 ? e % CT_RB : e / CT_KC, j = p.t_rs == 1 ? e / CT_RB : e % CT_KC;
        R tr = 0, ti = 0;
        if (i < nr && j < nk) {
          const long long off = (long long)(r0 + i) * p.t_rs + (long long)(k0 + j) * p.t_cs;
          tr = p.Ar[off]; ti = p.Ai[off];
        }
        Tr[j][i] = tr; Ti[j][i] = ti;
      }
      __syncthreads();
      if (active) {
        for (int j = 0; j < nk; j++) {
          const R yr = Vr[(long long)(k0 + j) * p.es], yi = Vi[(long long)(k0 + j) * p.es];
#pragma unroll
          for (int i = 0; i < CT_RB; i++) {
            const R tr = Tr[j][i], ti = Ti[j][i];
            const R pr = tr * yr - ti * yi, pi = tr * yi + ti * yr;
            ar[i] += SOLVE ? -pr : pr; ai[i] += SOLVE ? -pi : pi;
          }
        }
      }
    }
    // the chunk's own 16 x 16 triangle: T[j][i] = Teff(r0 + i, r0 + j); identity outside the block edge / on a unit diagonal
    __syncthreads();
    for (int e = tid; e < CT_RB * CT_RB; e += CT_THREADS) {
      const int i = p.t_rs == 1 ? e % CT_RB : e / CT_RB, j = p.t_rs == 1 ? e / CT_RB : e % CT_RB;
      R tr = (i == j) ? R(1) : R(0), ti = 0;
      const bool in = LOWER ? (j <= i) : (j >= i);
      if (i < nr && j < nr && in && !(p.unit && i == j)) {
        const long long off = (long long)(r0 + i) * p.t_rs + (long long)(r0 + j) * p.t_cs;
        tr = p.Ar[off]; ti = p.Ai[off];
      }
      Tr[j][i] = tr; Ti[j][i] = ti;
    }
    __syncthreads();
    if (active) {
      if (SOLVE) {
#pragma unroll
        for (int ii = 0; ii < CT_RB; ii++) {
          const int i = LOWER ? ii : CT_RB - 1 - ii;
          // x_i = (rhs_i - sum_{j solved} t_ij x_j) / t_ii   (complex division by the diagonal)
          R sr = ar[i], si = ai[i];
#pragma unroll
          for (int jj = 0; jj < CT_RB; jj++) {
            const bool dep = LOWER ? (jj < i) : (jj > i);
            if (dep) {
              const R tr = Tr[jj][i], ti = Ti[jj][i];
              sr -= tr * ar[jj] - ti * ai[jj];
              si -= tr * ai[jj] + ti * ar[jj];
            }
          }
          const R dr = Tr[i][i], di = Ti[i][i];
          const R den = dr * dr + di * di;
          ar[i] = (sr * dr + si * di) / den;
          ai[i] = (si * dr - sr * di) / den;
        }
      } else {
#pragma unroll
        for (int i = 0; i < CT_RB; i++)
#pragma unroll
          for (int j = 0; j < CT_RB; j++) {
            const bool in = LOWER ? (j <= i) : (j >= i);
            if (in) {
              const R tr = Tr[j][i], ti = Ti[j][i];
              ar[i] += tr * xr[j] - ti * xi[j];
              ai[i] += tr * xi[j] + ti * xr[j];
            }
          }
      }
#pragma unroll
      for (int i = 0; i < CT_RB; i++)
        if (i < nr) { Vr[(long long)(r0 + i) * p.es] = ar[i]; Vi[(long long)(r0 + i) * p.es] = ai[i]; }
    }
  }
}

}  // namespace nla

// tri_guard.cuh -- conditioning guard of the block-inverse solve leaves (Float32 / Float16) and the substitution they fall back to.
//
// A block-inverse leaf  X_blk = inv(Teff_blk) * V_blk  (tri_inv.cuh) is only conditionally backward stable: its residual grows with
// eps * cond(Teff_blk), where the reference's leaf -- substitution, src/trsm.jl:15-27 -- is backward stable for any triangular block,
// and the rounded inverse of a badly scaled block can overflow Float16 altogether.  So every inverted block gets a record
//     rec[0] = ||Teff_blk||_F^2     rec[1] = ||inv(Teff_blk) rounded to T||_F^2     rec[2] = number of non-finite entries of that inverse
// computed on the device right after the inverses (tri_cond_kernel), and BOTH consumers evaluate the same predicate on it:
//     bad  <=>  rec[2] > 0  or  ||T||_F ||inv T||_F / order  >  kappa_max          (= 1 for the identity; >= cond_2 / order)
// The leaf GEMM returns at once when its block is bad (GemmTcParams::skip_rec), and tri_subst_kernel -- launched right behind it --
// returns at once when it is NOT: exactly one of the two writes the block.  No host round trip, the call stays asynchronous; the price
// is one empty launch per inverted block (16 for n = 16384).  rec[3] counts the blocks that took the fallback.
//
// tri_subst_kernel: one thread per right-hand-side vector, rows in chunks of 16 held in registers, the chunk's row panel of Teff staged
// in shared memory and broadcast; FP32 arithmetic (as in the generic leaf, leaf.cuh).  Correct for every side / uplo / trans / diag and
// ragged sizes; slow (a few TFLOP/s) -- it only runs for blocks the guard rejects.
#pragma once
#include "common.cuh"

namespace nla {

__device__ __forceinline__ bool tri_cond_bad(const double* rec, double thr2) {
  const double r0 = rec[0], r1 = rec[1], r2 = rec[2];
  return r2 > 0.0 || !(r0 * r1 <= thr2);   // NaN in either norm counts as bad
}

template <typename T>
struct TriCondParams {
  const T* A; long long t_rs, t_cs;   // Teff(r,k) = A[r*t_rs + k*t_cs]
  int n, ib, lower, unit;
  int blk0;                           // first inverted block of this launch (block = blockIdx.y + blk0)
  const T* W;                         // rounded inverses, K-major, pitch ib (tri_inv.cuh)
  double* rec;                        // 4 doubles per block
};

template <typename T>
__global__ void __launch_bounds__(256) tri_cond_kernel(const TriCondParams<T> p) {
  const int b = blockIdx.y + p.blk0;
  const int off = b * p.ib;
  const int sz = min(p.ib, p.n - off);
  if (sz <= 0) return;
  double st = 0.0, sw = 0.0, bad = 0.0;
  const long long total = (long long)sz * sz;
  const bool r_fast = p.t_rs == 1;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int fast = (int)(e % sz), slow = (int)(e / sz);
    const int r = r_fast ? fast : slow, k = r_fast ? slow : fast;
    const bool in = p.lower ? (k <= r) : (k >= r);
    if (in) {
      const double v = (p.unit && r == k) ? 1.0 : (double)Traits<T>::ld(p.A + (long long)(off + r) * p.t_rs + (long long)(off + k) * p.t_cs);
      st += v * v;
    }
    // the inverse: row = slow, k = fast (K-major, contiguous in k)
    const double w = (double)Traits<T>::ld(p.W + (long long)(off + slow) * p.ib + fast);
    if (isfinite(w)) sw += w * w; else bad += 1.0;
  }
  __shared__ double red[3][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    st += __shfl_xor_sync(0xffffffffu, st, o);
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = st; red[1][warp] = sw; red[2][warp] = bad; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int w = 0; w < 8; w++) s += red[threadIdx.x][w];
    atomicAdd(p.rec + 4 * b + threadIdx.x, s);
  }
}

constexpr int TS_THREADS = 128, TS_RB = 16, TS_KC = 64;

template <typename T>
struct TriSubstParams {
  const T* A; long long t_rs, t_cs;   // Teff(r,k) strides
  int off, sz;                        // the diagonal block [off, off+sz)
  int unit;
  T* V; long long es, vs;             // element i of vector v at V[i*es + v*vs]
  int v0, nv;                         // vectors handled by this launch
  float scale;                        // alpha folded into this leaf (pre * post)
  const double* rec; double thr2;     // conditioning record of the block; the kernel runs only when it is bad
  double* counter;                    // += 1 when the fallback ran (one thread of the launch)
};

template <typename T, bool LOWER>
__global__ void __launch_bounds__(TS_THREADS) tri_subst_kernel(const TriSubstParams<T> p) {
  // launched with programmatic dependent launch: the successor's prologue may start at once, and this kernel waits here for the leaf
  // GEMM in front of it (whose own wait covered everything before), so that an unused fallback costs a few microseconds, not a drain
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (!tri_cond_bad(p.rec, p.thr2)) return;
  if (blockIdx.x == 0 && threadIdx.x == 0 && p.counter) atomicAdd(p.counter, 1.0);
  __shared__ __align__(16) float Ts[TS_KC][TS_RB];   // Ts[j][i] = Teff(r0 + i, k0 + j): one float4 x 4 per k for the 16 rows
  const int tid = threadIdx.x;
  const int v = p.v0 + blockIdx.x * TS_THREADS + tid;
  const bool active = v < p.v0 + p.nv;
  T* Vv = p.V + (long long)v * p.vs;
  const T* Ab = p.A + (long long)p.off * (p.t_rs + p.t_cs);
  const int nchunks = (p.sz + TS_RB - 1) / TS_RB;
  for (int c = 0; c < nchunks; c++) {
    const int r0 = (LOWER ? c : nchunks - 1 - c) * TS_RB;
    const int nr = min(TS_RB, p.sz - r0);
    float acc[TS_RB];
#pragma unroll
    for (int i = 0; i < TS_RB; i++)
      acc[i] = (active && i < nr) ? p.scale * (float)Traits<T>::ld(Vv + (long long)(p.off + r0 + i) * p.es) : 0.f;
    // already solved unknowns: k < r0 (lower) or k >= r0 + nr (upper)
    const int kbeg = LOWER ? 0 : r0 + nr, kend = LOWER ? r0 : p.sz;
    for (int k0 = kbeg; k0 < kend; k0 += TS_KC) {
      const int nk = min(TS_KC, kend - k0);
      __syncthreads();
      for (int e = tid; e < TS_KC * TS_RB; e += TS_THREADS) {
        // read along the contiguous direction of A
        const int i = p.t_rs == 1 ? e % TS_RB : e / TS_KC, j = p.t_rs == 1 ? e / TS_RB : e % TS_KC;
        float t = 0.f;
        if (i < nr && j < nk) t = (float)Traits<T>::ld(Ab + (long long)(r0 + i) * p.t_rs + (long long)(k0 + j) * p.t_cs);
        Ts[j][i] = t;
      }
      __syncthreads();
      if (active) {
        for (int j = 0; j < nk; j++) {
          const float x = (float)Traits<T>::ld(Vv + (long long)(p.off + k0 + j) * p.es);
          const float4* tp = reinterpret_cast<const float4*>(&Ts[j][0]);
#pragma unroll
          for (int q = 0; q < TS_RB / 4; q++) {
            const float4 t = tp[q];
            acc[4 * q + 0] = fmaf(-t.x, x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(-t.y, x, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(-t.z, x, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(-t.w, x, acc[4 * q + 3]);
          }
        }
      }
    }
    // the chunk's own 16 x 16 triangle: Ts[j][i] = Teff(r0 + i, r0 + j)
    __syncthreads();
    for (int e = tid; e < TS_RB * TS_RB; e += TS_THREADS) {
      const int i = p.t_rs == 1 ? e % TS_RB : e / TS_RB, j = p.t_rs == 1 ? e / TS_RB : e % TS_RB;
      float t = (i == j) ? 1.f : 0.f;
      const bool in = LOWER ? (j <= i) : (j >= i);
      if (i < nr && j < nr && in && !(p.unit && i == j)) t = (float)Traits<T>::ld(Ab + (long long)(r0 + i) * p.t_rs + (long long)(r0 + j) * p.t_cs);
      Ts[j][i] = t;
    }
    __syncthreads();
    if (active) {
      // substitution in the reference's scaled form (src/trsm.jl:15-27): x_i = b_i / d_i - sum (a_ij / d_i) x_j
#pragma unroll
      for (int ii = 0; ii < TS_RB; ii++) {
        const int i = LOWER ? ii : TS_RB - 1 - ii;
        const float rd = 1.f / Ts[i][i];
        float s = acc[i] * rd;
#pragma unroll
        for (int jj = 0; jj < TS_RB; jj++) {
          const bool dep = LOWER ? (jj < i) : (jj > i);
          if (dep) s = fmaf(-(Ts[jj][i] * rd), acc[jj], s);
        }
        acc[i] = s;
      }
#pragma unroll
      for (int i = 0; i < TS_RB; i++)
        if (i < nr) Traits<T>::st(Vv + (long long)(p.off + r0 + i) * p.es, acc[i]);
    }
  }
}

}  // namespace nla

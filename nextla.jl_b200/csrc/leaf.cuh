// leaf.cuh -- diagonal-block leaves: triangular solve / multiply of one t x t block (t <= 128) against a
// panel of right-hand-side vectors.
//
// Replaces the reference's four TRSM kernels (src/trsm.jl:5-126: one work-group per RHS vector, n threads,
// one barrier per pivot, A re-read from global memory for every vector) and four TRMM kernels
// (src/trmm.jl:43-312: 16x16 tiles) with ONE templated kernel:
//   * the block is normalised to "effective lower, forward" by index reversal at load time, and to
//     "vectors along the fast smem axis" whatever the side (left: vectors = columns of B, right: rows);
//   * the triangular tile is loaded ONCE per CTA into shared memory, already in the reference's scaled
//     form l'_rk = a_rk / d_r, b'_r = b_r / d_r (src/trsm.jl:15-18,24), so the substitution is division free;
//   * each thread owns one RHS vector; 16-row micro-blocks of the solution live in registers, the
//     already-solved part of the vector and the tile are read from shared memory (tile reads are
//     warp-uniform broadcasts, vector reads are conflict free), so there are NO barriers inside the solve
//     (the reference needs n of them).
// Round-1 kernel: CUDA-core FMA (FP64 DFMA peak equals the DMMA peak on B200, see profiles/).  The
// DMMA/TMA-staged variant is the planned next step.
#pragma once
#include "common.cuh"

namespace nla {

constexpr int LEAF_W = 128;        // RHS vectors per CTA == threads per CTA
constexpr int LEAF_MAX = 128;      // largest diagonal block one launch handles
constexpr int LEAF_PITCH = LEAF_W + 1;

template <typename T>
struct LeafParams {
  const T* A; long long a_rs, a_cs;  // Teff(r,k) = A[r*a_rs + k*a_cs] (block origin; already "transposed" by strides)
  int lower;                         // 1: Teff lower triangular, 0: upper
  int unit;                          // 1: unit diagonal -- the stored diagonal is not read (BLAS diag = 'U'; the reference's trsm/trmm
                                     // wrappers accept the flag and ignore it, src/trsm.jl:186, src/trmm.jl:430)
  int t;                             // block order
  T* V; long long es, vs;            // vector v, element e at V[e*es + v*vs]
  long long m;                       // number of vectors
  double pre, post;                  // b <- pre*b before a solve (src/rectrxm.jl:64); result <- post*result after a multiply (:72)
};

template <typename T>
inline size_t leaf_smem_bytes(int t) {
  using Acc = typename Traits<T>::Acc;
  const int tp = (t + 15) & ~15, nrb = tp / 16;
  return sizeof(Acc) * ((size_t)128 * nrb * (nrb + 1) + tp + (size_t)tp * LEAF_PITCH);
}

template <typename T, bool SOLVE>
__global__ void __launch_bounds__(LEAF_W) leaf_kernel(const LeafParams<T> p) {
  using Acc = typename Traits<T>::Acc;
  extern __shared__ __align__(16) uint8_t leaf_smem[];
  const int t = p.t, tp = (t + 15) & ~15, nrb = tp / 16;
  Acc* Ls = reinterpret_cast<Acc*>(leaf_smem);          // packed micro-block rows: block rb holds [k = 0 .. 16rb+15][16 rows]
  Acc* dv = Ls + 128 * nrb * (nrb + 1);                 // diagonal (normalised order)
  Acc* panel = dv + tp;                                 // [element][vector], pitch LEAF_PITCH
  const int tid = threadIdx.x;

  // ---- stage the triangular tile (normalised to lower/forward by index reversal for upper) ----
  for (int r = tid; r < tp; r += LEAF_W) {
    Acc d = Acc(1);
    if (r < t && !p.unit) {
      const int R = p.lower ? r : t - 1 - r;
      d = Traits<T>::ld(p.A + (long long)R * (p.a_rs + p.a_cs));
    }
    dv[r] = d;
  }
  __syncthreads();
  for (int rb = 0; rb < nrb; rb++) {
    Acc* Lb = Ls + 128 * rb * (rb + 1);
    const int cnt = (rb + 1) * 256;
    for (int e = tid; e < cnt; e += LEAF_W) {
      const int k = e >> 4, r = rb * 16 + (e & 15);
      Acc v = Acc(0);
      const bool inside = SOLVE ? (k < r) : (k <= r);
      if (inside && r < t) {
        const int R = p.lower ? r : t - 1 - r, K = p.lower ? k : t - 1 - k;
        v = (p.unit && k == r) ? Acc(1) : Traits<T>::ld(p.A + (long long)R * p.a_rs + (long long)K * p.a_cs);
        if (SOLVE) v = v / dv[r];   // the reference divides every entry by its row's diagonal (src/trsm.jl:24)
      }
      Lb[e] = v;
    }
  }

  // ---- stage the RHS panel: panel[e'][v] ----
  const long long v0 = (long long)blockIdx.x * LEAF_W;
  const int nv = (int)min((long long)LEAF_W, p.m - v0);
  T* Vb = p.V + v0 * p.vs;
  const Acc pre = (Acc)p.pre;
  if (p.es == 1) {  // elements contiguous (left side): a warp reads one vector, lanes along the elements
    const int warp = tid >> 5, lane = tid & 31;
    for (int v = warp; v < LEAF_W; v += LEAF_W / 32)
      for (int e = lane; e < tp; e += 32) {
        Acc b = Acc(0);
        if (v < nv && e < t) {
          const int E = p.lower ? e : t - 1 - e;
          b = Traits<T>::ld(Vb + (long long)v * p.vs + E);
          if (SOLVE) {
            if (p.pre != 1.0) { T tmp; Traits<T>::st(&tmp, pre * b); b = Traits<T>::ld(&tmp); }
            b = b / dv[e];
          }
        }
        panel[e * LEAF_PITCH + v] = b;
      }
  } else {  // vectors contiguous (right side): threads along the vectors
    for (int e = 0; e < tp; e++) {
      Acc b = Acc(0);
      if (tid < nv && e < t) {
        const int E = p.lower ? e : t - 1 - e;
        b = Traits<T>::ld(Vb + (long long)tid * p.vs + (long long)E * p.es);
        if (SOLVE) {
          if (p.pre != 1.0) { T tmp; Traits<T>::st(&tmp, pre * b); b = Traits<T>::ld(&tmp); }
          b = b / dv[e];
        }
      }
      panel[e * LEAF_PITCH + tid] = b;
    }
  }
  __syncthreads();

  // ---- one thread per vector: register micro-blocks of 16 rows, left-looking ----
  Acc* col = panel + tid;
  if (SOLVE) {
    for (int rb = 0; rb < nrb; rb++) {
      const Acc* Lb = Ls + 128 * rb * (rb + 1);
      Acc acc[16];
#pragma unroll
      for (int r = 0; r < 16; r++) acc[r] = col[(rb * 16 + r) * LEAF_PITCH];
      const int kend = rb * 16;
#pragma unroll 2
      for (int k = 0; k < kend; k++) {
        const Acc xk = col[k * LEAF_PITCH];
#pragma unroll
        for (int r = 0; r < 16; r++) acc[r] -= Lb[k * 16 + r] * xk;
      }
      const Acc* Ld = Lb + kend * 16;
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const Acc xi = acc[i];
#pragma unroll
        for (int r = i + 1; r < 16; r++) acc[r] -= Ld[i * 16 + r] * xi;
      }
#pragma unroll
      for (int r = 0; r < 16; r++) col[(rb * 16 + r) * LEAF_PITCH] = acc[r];
    }
  } else {
    for (int rb = nrb - 1; rb >= 0; rb--) {
      const Acc* Lb = Ls + 128 * rb * (rb + 1);
      Acc acc[16];
#pragma unroll
      for (int r = 0; r < 16; r++) acc[r] = Acc(0);
      const int kend = rb * 16;
#pragma unroll 2
      for (int k = 0; k < kend; k++) {
        const Acc bk = col[k * LEAF_PITCH];
#pragma unroll
        for (int r = 0; r < 16; r++) acc[r] += Lb[k * 16 + r] * bk;
      }
      const Acc* Ld = Lb + kend * 16;
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const Acc bi = col[(kend + i) * LEAF_PITCH];
#pragma unroll
        for (int r = i; r < 16; r++) acc[r] += Ld[i * 16 + r] * bi;
      }
#pragma unroll
      for (int r = 0; r < 16; r++) col[(rb * 16 + r) * LEAF_PITCH] = acc[r];
    }
  }
  __syncthreads();

  // ---- write the panel back ----
  const Acc post = (Acc)p.post;
  if (p.es == 1) {
    const int warp = tid >> 5, lane = tid & 31;
    for (int v = warp; v < nv; v += LEAF_W / 32)
      for (int e = lane; e < t; e += 32) {
        const int E = p.lower ? e : t - 1 - e;
        Acc x = panel[e * LEAF_PITCH + v];
        T* dst = Vb + (long long)v * p.vs + E;
        if (!SOLVE && p.post != 1.0) { Traits<T>::st(dst, x); x = post * Traits<T>::ld(dst); }
        Traits<T>::st(dst, x);
      }
  } else if (tid < nv) {
    for (int e = 0; e < t; e++) {
      const int E = p.lower ? e : t - 1 - e;
      Acc x = panel[e * LEAF_PITCH + tid];
      T* dst = Vb + (long long)tid * p.vs + (long long)E * p.es;
      if (!SOLVE && p.post != 1.0) { Traits<T>::st(dst, x); x = post * Traits<T>::ld(dst); }
      Traits<T>::st(dst, x);
    }
  }
}

}  // namespace nla

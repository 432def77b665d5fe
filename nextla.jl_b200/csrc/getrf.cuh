// getrf.cuh -- the panel factorisation of the reference's recursive LU, `getrf2!` (src/lu.jl:185-299), on the device.
//
// The reference recurses on the column count down to single columns (:216-251: pivot search by |re|+|im|, first maximum wins; the
// interchange; the column scaled by 1/pivot, divided instead when |pivot| < sfmin; a zero pivot recorded in info and the column left
// alone).  Here the host recursion (nla_api.cu, getrf2_rec) stops at panels of at most GETRF_NB columns, and one COOPERATIVE launch of
// getrf_panel_kernel factors a whole m x nb panel with the same pivot rule:
//   * the panel's rows are dealt to the CTAs in contiguous chunks and live in shared memory for the whole launch (one read and one
//     write of the panel in HBM),
//   * per column: every CTA finds its own candidate (largest |a|, lowest row on ties) and publishes it WITH that row's nb entries;
//     the CTA that owns row j also publishes row j; one grid barrier; every CTA then reduces the candidates redundantly, so it knows
//     the pivot row's contents (= row j of U, needed for its own rank-1 update) and the two owners store the exchanged rows.
//     The exchange area carries a sequence number in every word (see getrf_ll_words), so there is no barrier object at all: a column
//     costs one store and two polled reads by the first warp of each CTA.  (Measured per column: 5.2 us with cooperative_groups'
//     grid.sync() and two block-wide read phases behind it, 7.7 us with a release/acquire counter; the launch stays cooperative for
//     the co-residency guarantee the polling needs.)  Buffers alternate with the column parity: a CTA can only publish column j + 2
//     after it has seen every CTA's column j + 1, i.e. after every CTA has finished reading column j.
//   * ipiv is written relative to the panel's first row (1-based, as the reference's view-relative pivots, :237), info with the
//     panel's column offset added (:260-262, :289-291) unless an earlier column already set it.
#pragma once
#include "common.cuh"

namespace nla {

constexpr int GETRF_NB = 64;        // widest panel
constexpr int GETRF_THREADS = 256;
constexpr int GETRF_UG = 16;         // columns per register group of the rank-1 update
constexpr int GETRF_SLOTS = 8;       // candidate headers polled per lane: up to 256 CTAs

// Exchange area of the panel kernel: every 8-byte word carries 32 bits of payload and the 32-bit sequence number of the column it
// belongs to, so a reader needs no flag, counter or fence -- it polls the words it wants until they show the right sequence number
// (8-byte accesses are single-copy atomic).  A double travels as two words.  Per column parity: G headers of 4 words
// {magnitude lo, magnitude hi, row, unused}, G candidate rows of GETRF_NB doubles, and row j before the interchange.
constexpr size_t getrf_ll_words(int G) { return 2 * ((size_t)G * 4 + (size_t)G * GETRF_NB * 2 + (size_t)GETRF_NB * 2); }

template <typename T>
struct GetrfPanelParams {
  T* A; long long lda;
  int m, n;                // panel: m rows, n <= GETRF_NB columns
  int rows_per_cta;
  long long* ipiv;         // min(m, n) entries
  int* info;
  int col_off;             // 0-based global column of the panel's first column
  unsigned long long* ll;  // exchange area, getrf_ll_words(gridDim.x) words (zeroed once when allocated)
  unsigned seq_base;       // column j of this launch uses sequence number seq_base + j + 1 (unique across launches on the handle)
  double sfmin;
};

__device__ __forceinline__ void ll_put_double(unsigned long long* w, double v, unsigned seq) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v), s = (unsigned long long)seq << 32;
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(w), "l"(s | (b & 0xffffffffull)), "l"(s | (b >> 32)) : "memory");
}
__device__ __forceinline__ void ll_peek2(const unsigned long long* w, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(w) : "memory");
}
__device__ __forceinline__ void ll_peek3(const unsigned long long* w, unsigned long long& a, unsigned long long& b, unsigned long long& d) {
  unsigned long long pad;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(w) : "memory");
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(d), "=l"(pad) : "l"(w + 2) : "memory");
}
__device__ __forceinline__ void ll_put_int(unsigned long long* w, int v, unsigned seq) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(w), "l"(((unsigned long long)seq << 32) | (unsigned)v) : "memory");
}
// Warp-wide "largest magnitude, lowest row on ties" on (best, bi) pairs; best >= 0 (or -1 with bi = INT_MAX for "no candidate").  The bit
// pattern of a non-negative double orders like an unsigned integer, so three redux.sync instructions (high word, low word among the
// lanes that match the high word, lowest row among the lanes that match both) replace five rounds of 64-bit shuffles and FP64 compares
// (measured with clock64 stamps: ~100 cycles per shuffle round, three such reductions per column).  Every lane gets the result.
__device__ __forceinline__ void getrf_warp_argmax(double& best, int& bi) {
  const unsigned long long key = bi == 0x7fffffff ? 0ull : (unsigned long long)__double_as_longlong(best);
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  const unsigned mi = __reduce_min_sync(0xffffffffu, (hi == mh && lo == ml) ? (unsigned)bi : 0x7fffffffu);
  bi = (int)mi;
  best = mi == 0x7fffffffu ? -1.0 : __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
}

// Row i of the rank-1 update of step j: l = a_ij / pivot, a_ik -= l * u_jk for k > j; returns |a_i,j+1| after the update (the row's
// candidate for the next column; -1 when there is none).  The columns are taken in groups of GETRF_UG with all loads of a group ahead of its
// stores: S and the pivot row are both in shared memory, and with one load - FMA - store per column the compiler must keep every load
// behind the previous store (possible alias), which made a column cost ~60 cycles of latency each -- 2 us per step whatever the row count.
template <typename T>
__device__ __forceinline__ double getrf_update_row(T* __restrict__ Si, int rp, int j, int n, bool mul, T inv, T piv, const T* __restrict__ prow) {
  T l = Si[j * rp];
  l = mul ? l * inv : l / piv;
  Si[j * rp] = l;
  double next = -1.0;
  int k = j + 1;
  if (k < n) {
    T x[GETRF_UG];
#pragma unroll
    for (int u = 0; u < GETRF_UG; u++) x[u] = (k + u < n) ? Si[(k + u) * rp] : T(0);
#pragma unroll
    for (int u = 0; u < GETRF_UG; u++) x[u] -= l * ((k + u < n) ? prow[k + u] : T(0));
#pragma unroll
    for (int u = 0; u < GETRF_UG; u++) if (k + u < n) Si[(k + u) * rp] = x[u];
    next = fabs((double)x[0]);
    if (next != next) next = __longlong_as_double(0x7ff0000000000000ll);
    k += GETRF_UG;
  }
  for (; k < n; k += GETRF_UG) {
    T x[GETRF_UG];
#pragma unroll
    for (int u = 0; u < GETRF_UG; u++) x[u] = (k + u < n) ? Si[(k + u) * rp] : T(0);
#pragma unroll
    for (int u = 0; u < GETRF_UG; u++) x[u] -= l * ((k + u < n) ? prow[k + u] : T(0));
#pragma unroll
    for (int u = 0; u < GETRF_UG; u++) if (k + u < n) Si[(k + u) * rp] = x[u];
  }
  return next;
}

template <typename T>
__global__ void __launch_bounds__(GETRF_THREADS) getrf_panel_kernel(const GetrfPanelParams<T> p) {
  extern __shared__ __align__(16) unsigned char getrf_smem[];
  T* S = reinterpret_cast<T*>(getrf_smem);                 // S[k * rp + i]: column k, local row i
  __shared__ T prow[GETRF_NB], orow[GETRF_NB];
  __shared__ double wv[GETRF_THREADS / 32];
  __shared__ int wi[GETRF_THREADS / 32];
  __shared__ int s_row;
  const int G = gridDim.x, c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rp = p.rows_per_cta;
  const int r0 = c * rp, nr = min(p.m, r0 + rp) - r0;
  // panel -> shared memory, 8 columns of a row in flight per thread (one column at a time waits a memory round trip per column:
  // ncu source view, 13 % of the cluster kernel's samples sat on the store behind that load)
  for (int i = tid; i < nr; i += GETRF_THREADS)
    for (int k0 = 0; k0 < p.n; k0 += 8) {
      T v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = (k0 + u < p.n) ? p.A[(long long)(k0 + u) * p.lda + r0 + i] : T(0);
#pragma unroll
      for (int u = 0; u < 8; u++) if (k0 + u < p.n) S[(k0 + u) * rp + i] = v[u];
    }
  __syncthreads();
  const int steps = min(p.m, p.n);
  bool have = false; double nbest = -1.0; int nbi = 0x7fffffff;   // candidate carried over from the previous column's update
  for (int j = 0; j < steps; j++) {
    const int buf = j & 1;
    // this CTA's candidate: first row of the largest magnitude among its rows >= j (a NaN counts as the largest)
    double best = nbest; int bi = nbi;
    if (!have) {
      best = -1.0; bi = 0x7fffffff;
      for (int i = tid; i < nr; i += GETRF_THREADS) {
        if (r0 + i < j) continue;
        double v = fabs((double)S[j * rp + i]);
        if (v != v) v = __longlong_as_double(0x7ff0000000000000ll);
        if (v > best) { best = v; bi = r0 + i; }
      }
    }
    getrf_warp_argmax(best, bi);
    if (lane == 0) { wv[warp] = best; wi[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      // the first warp finishes the CTA's candidate, publishes it with the candidate row's entries, collects everybody's candidates and
      // fetches the pivot row; the other warps wait at the next __syncthreads
      best = lane < GETRF_THREADS / 32 ? wv[lane] : -1.0;
      bi = lane < GETRF_THREADS / 32 ? wi[lane] : 0x7fffffff;
      getrf_warp_argmax(best, bi);
      best = __shfl_sync(0xffffffffu, best, 0); bi = __shfl_sync(0xffffffffu, bi, 0);
      const unsigned seq = p.seq_base + (unsigned)j + 1u;
      unsigned long long* hdr = p.ll + (size_t)buf * ((size_t)G * 4 + (size_t)G * GETRF_NB * 2 + (size_t)GETRF_NB * 2);
      unsigned long long* rows = hdr + (size_t)G * 4;
      unsigned long long* rowj = rows + (size_t)G * GETRF_NB * 2;
      if (lane == 0) ll_put_double(hdr + (size_t)c * 4, best, seq);
      if (lane == 1) ll_put_int(hdr + (size_t)c * 4 + 2, bi, seq);
      if (bi != 0x7fffffff)
        for (int k = lane; k < p.n; k += 32) ll_put_double(rows + ((size_t)c * GETRF_NB + k) * 2, (double)S[k * rp + (bi - r0)], seq);
      if (j >= r0 && j < r0 + nr)                      // the owner of row j publishes it as it is before the interchange
        for (int k = lane; k < p.n; k += 32) ll_put_double(rowj + (size_t)k * 2, (double)S[k * rp + (j - r0)], seq);
      // all of a lane's polls are in flight together: a pass issues every outstanding load, then looks at what came back
      double v = -1.0; int r = 0x7fffffff;
      {
        unsigned pending = 0;
#pragma unroll
        for (int sl = 0; sl < GETRF_SLOTS; sl++) if (lane + 32 * sl < G) pending |= 1u << sl;
        while (pending) {
          unsigned long long a[GETRF_SLOTS], b[GETRF_SLOTS], d[GETRF_SLOTS];
#pragma unroll
          for (int sl = 0; sl < GETRF_SLOTS; sl++)
            if (pending >> sl & 1) ll_peek3(hdr + (size_t)(lane + 32 * sl) * 4, a[sl], b[sl], d[sl]);
#pragma unroll
          for (int sl = 0; sl < GETRF_SLOTS; sl++)
            if ((pending >> sl & 1) && (unsigned)(a[sl] >> 32) == seq && (unsigned)(b[sl] >> 32) == seq && (unsigned)(d[sl] >> 32) == seq) {
              const double gv = __longlong_as_double((long long)((a[sl] & 0xffffffffull) | (b[sl] << 32)));
              const int gr = (int)(unsigned)d[sl];
              if (gv > v || (gv == v && gr < r)) { v = gv; r = gr; }
              pending &= ~(1u << sl);
            }
        }
      }
      getrf_warp_argmax(v, r);
      const bool mine = r >= r0 && r < r0 + nr && r != j;     // this CTA stores the old row j at the pivot's position
      {
        const unsigned long long* wrow = rows + (size_t)(r / rp) * GETRF_NB * 2;
        unsigned pending = 0;    // bits 0,1: pivot row entries lane, lane + 32; bits 2,3: the same entries of row j
        if (lane < p.n) pending |= mine ? 5u : 1u;
        if (lane + 32 < p.n) pending |= mine ? 10u : 2u;
        while (pending) {
          unsigned long long a[4], b[4];
#pragma unroll
          for (int sl = 0; sl < 4; sl++)
            if (pending >> sl & 1) ll_peek2((sl < 2 ? wrow : rowj) + (size_t)(lane + 32 * (sl & 1)) * 2, a[sl], b[sl]);
#pragma unroll
          for (int sl = 0; sl < 4; sl++)
            if ((pending >> sl & 1) && (unsigned)(a[sl] >> 32) == seq && (unsigned)(b[sl] >> 32) == seq) {
              const T x = (T)__longlong_as_double((long long)((a[sl] & 0xffffffffull) | (b[sl] << 32)));
              (sl < 2 ? prow : orow)[lane + 32 * (sl & 1)] = x;
              pending &= ~(1u << sl);
            }
        }
      }
      if (lane == 0) s_row = r;
    }
    __syncthreads();
    const int pr = s_row;
    const T piv = prow[j];
    if (c == 0 && tid == 0) {
      p.ipiv[j] = (long long)pr + 1;
      if (piv == T(0) && *p.info == 0) *p.info = p.col_off + j + 1;
    }
    if (pr != j && tid < p.n) {
      if (j >= r0 && j < r0 + nr) S[tid * rp + (j - r0)] = prow[tid];
      if (pr >= r0 && pr < r0 + nr) S[tid * rp + (pr - r0)] = orow[tid];
    }
    __syncthreads();
    // scale column j, rank-1 update of the columns to its right; the thread's candidate for column j + 1 falls out of the update
    have = (piv != T(0)) && (j + 1 < p.n);
    nbest = -1.0; nbi = 0x7fffffff;
    if (piv != T(0)) {
      const bool mul = fabs((double)piv) >= p.sfmin;
      const T inv = T(1) / piv;
      for (int i = tid; i < nr; i += GETRF_THREADS) {
        if (r0 + i <= j) continue;
        const double a = getrf_update_row<T>(S + i, rp, j, p.n, mul, inv, piv, prow);
        if (a > nbest) { nbest = a; nbi = r0 + i; }
      }
    }
    // no barrier here: the next column's __syncthreads orders these writes before the first warp reads S, and prow / orow / wv / wi are
    // rewritten only behind it
  }
  __syncthreads();
  for (int k = 0; k < p.n; k++)
    for (int i = tid; i < nr; i += GETRF_THREADS) p.A[(long long)k * p.lda + r0 + i] = S[k * rp + i];
}

// ---- panel kernel, one thread-block cluster -------------------------------------------------------------------------------------
// Same algorithm and pivot rule as getrf_panel_kernel, for panels whose rows fit the shared memory of ONE cluster (8 CTAs, 16 with the
// non-portable size): the exchange goes through distributed shared memory instead of L2.  Per column every CTA PUSHES its candidate
// (magnitude, row, the row's entries) into a slot of every CTA of the cluster, the owner of row j pushes row j, one hardware cluster
// barrier (barrier.cluster arrive.release / wait.acquire) makes the pushes visible, and from there on every CTA works on local copies:
// it reduces the candidates itself and has row j of U for its rank-1 update.  Slots alternate with the column parity (a CTA can only
// push column j + 2 after the barrier of column j + 1, which every CTA reaches after it is done with column j's slots).
// Two __syncthreads and one cluster barrier per column instead of two L2 round trips.
constexpr int GETRF_CL_MAX = 16;

template <typename T>
struct GetrfClusterParams {
  T* A; long long lda;
  int m, n;
  int rows_per_cta;
  long long* ipiv;
  int* info;
  int col_off;
  double sfmin;
  long long* dbg;          // probes only: per-phase clock totals of CTA 0 (null = off)
};

__device__ __forceinline__ void st_cluster_b64(uint32_t addr, unsigned long long v) {
  asm volatile("st.shared::cluster.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_b32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
template <typename T> __device__ __forceinline__ void st_cluster_val(uint32_t addr, T v);
template <> __device__ __forceinline__ void st_cluster_val<double>(uint32_t addr, double v) { st_cluster_b64(addr, (unsigned long long)__double_as_longlong(v)); }
template <> __device__ __forceinline__ void st_cluster_val<float>(uint32_t addr, float v) { st_cluster_b32(addr, __float_as_uint(v)); }

template <typename T>
__global__ void __launch_bounds__(GETRF_THREADS) getrf_panel_cluster_kernel(const GetrfClusterParams<T> p) {
  extern __shared__ __align__(16) unsigned char getrf_smem[];
  T* S = reinterpret_cast<T*>(getrf_smem);                 // S[k * rp + i]: column k, local row i
  __shared__ T rows_s[2][GETRF_CL_MAX][GETRF_NB];          // candidate rows of every CTA of the cluster
  __shared__ T rowj_s[2][GETRF_NB];                        // row j before the interchange
  __shared__ double cval_s[2][GETRF_CL_MAX];
  __shared__ int cidx_s[2][GETRF_CL_MAX];
  __shared__ double wv[GETRF_THREADS / 32];
  __shared__ int wi[GETRF_THREADS / 32];
  const int CL = gridDim.x, c = (int)cluster_ctarank(), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rp = p.rows_per_cta;
  const int r0 = c * rp, nr = max(0, min(p.m, r0 + rp) - r0);
  // panel -> shared memory, 8 columns of a row in flight per thread (one column at a time waits a memory round trip per column:
  // ncu source view, 13 % of the cluster kernel's samples sat on the store behind that load)
  for (int i = tid; i < nr; i += GETRF_THREADS)
    for (int k0 = 0; k0 < p.n; k0 += 8) {
      T v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = (k0 + u < p.n) ? p.A[(long long)(k0 + u) * p.lda + r0 + i] : T(0);
#pragma unroll
      for (int u = 0; u < 8; u++) if (k0 + u < p.n) S[(k0 + u) * rp + i] = v[u];
    }
  cluster_sync_all();                                      // every CTA of the cluster is running before the first remote store
  const int steps = min(p.m, p.n);
  bool have = false; double nbest = -1.0; int nbi = 0x7fffffff;
  long long ph[6] = {0, 0, 0, 0, 0, 0}, tprev = clock64();
#define GETRF_PH(x) do { if (p.dbg) { const long long tn = clock64(); ph[x] += tn - tprev; tprev = tn; } } while (0)
  for (int j = 0; j < steps; j++) {
    const int par = j & 1;
    double best = nbest; int bi = nbi;
    if (!have) {
      best = -1.0; bi = 0x7fffffff;
      for (int i = tid; i < nr; i += GETRF_THREADS) {
        if (r0 + i < j) continue;
        double v = fabs((double)S[j * rp + i]);
        if (v != v) v = __longlong_as_double(0x7ff0000000000000ll);
        if (v > best) { best = v; bi = r0 + i; }
      }
    }
    getrf_warp_argmax(best, bi);
    if (lane == 0) { wv[warp] = best; wi[warp] = bi; }
    GETRF_PH(0);
    __syncthreads();                                       // also: the previous column's update of S is complete
    GETRF_PH(1);
    // every warp finishes the CTA's candidate on its own (8 partial results, one per lane, three exchange rounds)
    best = wv[lane & (GETRF_THREADS / 32 - 1)]; bi = wi[lane & (GETRF_THREADS / 32 - 1)];
    getrf_warp_argmax(best, bi);
    // push: header to slot c of every CTA, candidate row and (owner only) row j; thread (k, r) serves entry k for the CTAs r, r + 4, ...
    if (tid < CL) {
      st_cluster_b64(mapa_u32(smem_u32(&cval_s[par][c]), (uint32_t)tid), (unsigned long long)__double_as_longlong(best));
      st_cluster_b32(mapa_u32(smem_u32(&cidx_s[par][c]), (uint32_t)tid), (uint32_t)bi);
    }
    {
      const int k = tid & (GETRF_NB - 1);
      const bool own_j = j >= r0 && j < r0 + nr;
      if (k < p.n && (bi != 0x7fffffff || own_j)) {
        const T cand = bi != 0x7fffffff ? S[k * rp + (bi - r0)] : T(0);
        const T rj = own_j ? S[k * rp + (j - r0)] : T(0);
        const uint32_t a_rows = smem_u32(&rows_s[par][c][k]), a_rowj = smem_u32(&rowj_s[par][k]);
        for (int r = tid / GETRF_NB; r < CL; r += GETRF_THREADS / GETRF_NB) {
          if (bi != 0x7fffffff) st_cluster_val<T>(mapa_u32(a_rows, (uint32_t)r), cand);
          if (own_j) st_cluster_val<T>(mapa_u32(a_rowj, (uint32_t)r), rj);
        }
      }
    }
    GETRF_PH(2);
    cluster_sync_all();
    GETRF_PH(3);
    // every warp reduces the CL candidates from the local slots (one per lane, four exchange rounds)
    double v = -1.0; int pr = 0x7fffffff;
    if ((lane & (GETRF_CL_MAX - 1)) < CL) { v = cval_s[par][lane & (GETRF_CL_MAX - 1)]; pr = cidx_s[par][lane & (GETRF_CL_MAX - 1)]; }
    getrf_warp_argmax(v, pr);
    const T* prow = rows_s[par][pr / rp];
    const T piv = prow[j];
    if (c == 0 && tid == 0) {
      p.ipiv[j] = (long long)pr + 1;
      if (piv == T(0) && *p.info == 0) *p.info = p.col_off + j + 1;
    }
    if (pr != j && tid < p.n) {
      if (j >= r0 && j < r0 + nr) S[tid * rp + (j - r0)] = prow[tid];
      if (pr >= r0 && pr < r0 + nr) S[tid * rp + (pr - r0)] = rowj_s[par][tid];
    }
    __syncthreads();
    GETRF_PH(4);
    have = (piv != T(0)) && (j + 1 < p.n);
    nbest = -1.0; nbi = 0x7fffffff;
    if (piv != T(0)) {
      const bool mul = fabs((double)piv) >= p.sfmin;
      const T inv = T(1) / piv;
      for (int i = tid; i < nr; i += GETRF_THREADS) {
        if (r0 + i <= j) continue;
        const double a = getrf_update_row<T>(S + i, rp, j, p.n, mul, inv, piv, prow);
        if (a > nbest) { nbest = a; nbi = r0 + i; }
      }
    }
    GETRF_PH(5);
  }
#undef GETRF_PH
  if (p.dbg && c == 0 && tid == 0) for (int x = 0; x < 6; x++) p.dbg[x] = ph[x];
  cluster_sync_all();                                      // nobody exits while a neighbour may still push into its slots
  for (int k = 0; k < p.n; k++)
    for (int i = tid; i < nr; i += GETRF_THREADS) p.A[(long long)k * p.lda + r0 + i] = S[k * rp + i];
}

// ipiv[i] += shift for i < count: the reference's pivot adjustment after the second recursive call (src/lu.jl:293-295)
__global__ void __launch_bounds__(256) ipiv_shift_kernel(long long* __restrict__ ipiv, long long count, long long shift) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) ipiv[i] += shift;
}

}  // namespace nla

// slab_f64.cuh -- fused FP64 "macro-leaf": one CTA solves / multiplies a whole T x T diagonal block (T up to the
// full matrix) against its own slab of 128 right-hand-side vectors, left-looking, with NO inter-CTA communication.
//
// This is the B200 replacement for the bottom of the reference's recursion: its leaf kernels (src/trsm.jl:5-126,
// src/trmm.jl:43-312) plus the small-K GEMM_SUB!/GEMM_ADD! levels just above them (src/rectrxm.jl:159-197), which on
// B200 are launch- and tail-bound (profiles/r01_launches_step_summary.csv: leaves 1.9 TFLOP/s, K<=512 updates 10-21).
//
// Per block row i (128 rows) of the diagonal block:
//   1. S = sum_j Teff[i,j] * X[j]   over the already-final block rows j  -- DMMA main loop, operands staged by TMA
//      through a 4-stage mbarrier ring exactly as in gemm_f64.cuh (same swizzle / k-permutation);
//   2. (solve) rhs = beta*B[i] - S, then the 128x128 triangular solve IN REGISTERS: every warp owns all 128 rows of
//      16 vectors (accumulator fragments), 8-row micro-blocks are solved by warp-shuffle forward/back substitution
//      in the reference's scaled form l'_rk = a_rk/d_r, x_r = b_r/d_r - sum l'_rk x_k (src/trsm.jl:15-27), and the rows
//      below/above are updated with DMMAs whose A fragments (the diagonal tile of A) come straight from L1/L2;
//      (multiply) the diagonal block is just one more K block whose A fragments are masked to the triangle;
//   3. the block row is written back in place; a proxy fence + mbarrier tells the TMA producer that the rows it is
//      about to re-read as the next operand are in memory.
// The producer warp prefetches the non-dependent K blocks of row i+1 while the consumers are in step 2.
#pragma once
#include "gemm_f64.cuh"

namespace nla {

constexpr int SL_BM = 128, SL_W = 128, SL_BK = 16, SL_STAGES = 4;
constexpr int SL_TILE_BYTES = SL_BM * SL_BK * 8;
constexpr int SL_STAGE_BYTES = 2 * SL_TILE_BYTES;
constexpr int SL_CONSUMER_WARPS = 8;
constexpr int SL_THREADS = (SL_CONSUMER_WARPS + 4) * 32;  // 2 consumer warpgroups + 1 producer warpgroup (1 active warp)
constexpr int SL_SCRATCH_BYTES = SL_CONSUMER_WARPS * 8 * 16 * 8;
constexpr int SL_SMEM_BYTES = SL_STAGES * SL_STAGE_BYTES + SL_SCRATCH_BYTES + 1024;

struct SlabParams {
  int T;                  // order of the diagonal block
  int off;                // its origin (row == column) in Teff coordinates
  int v_base, v_count;    // vectors (columns of B) handled by this launch
  const double* A;        // Teff(r,c) = A[r*t_rs + c*t_cs]
  long long t_rs, t_cs;
  double* B;              // column-major, vectors are columns
  long long ldb;
  double beta, post;
  int unit;               // 1: unit diagonal (the stored diagonal is not used)
};

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int AMAJ, bool LOWER, bool SOLVE>
__global__ void __launch_bounds__(SL_THREADS, 1)
slab_f64_kernel(const __grid_constant__ CUtensorMap mapT, const __grid_constant__ CUtensorMap mapV, const SlabParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[SL_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[SL_STAGES];
  __shared__ __align__(8) uint64_t xdone_bar;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SL_STAGES; s++) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), SL_CONSUMER_WARPS);
    }
    mbar_init(smem_u32(&xdone_bar), SL_CONSUMER_WARPS);
    mbar_fence_init();
  }
  __syncthreads();

  const int nb = (p.T + SL_BM - 1) / SL_BM;
  constexpr bool ASC = (SOLVE == LOWER);   // block rows ascending: solve-lower and multiply-upper
  const int v0 = p.v_base + blockIdx.x * SL_W;          // first vector of this CTA (absolute column of B)
  const int v_end = p.v_base + p.v_count;

  // K-block sequence of block row i:  count and j(jj).  The block written by the previous row comes LAST.
  auto kcount = [&](int i) { return SOLVE ? (LOWER ? i : nb - 1 - i) : (LOWER ? i + 1 : nb - i); };
  auto kblock = [&](int i, int jj) { return SOLVE ? (LOWER ? jj : nb - 1 - jj) : (LOWER ? jj : i + jj); };
  auto ktiles = [&](int j) { return (min(SL_BM, p.T - j * SL_BM) + SL_BK - 1) / SL_BK; };

  if (warp >= SL_CONSUMER_WARPS) {
    // ===== TMA producer warpgroup: hands its registers to the consumers, one lane of one warp issues the copies =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == SL_CONSUMER_WARPS && lane == 0) {
      tma_prefetch_desc(&mapT);
      tma_prefetch_desc(&mapV);
      int kt = 0;
      for (int r = 0; r < nb; r++) {
        const int i = ASC ? r : nb - 1 - r;
        const int cnt = kcount(i);
        for (int jj = 0; jj < cnt; jj++) {
          const int j = kblock(i, jj);
          if (SOLVE && jj == cnt - 1) mbar_wait(smem_u32(&xdone_bar), (r - 1) & 1);  // rows written by the previous block row
          const int nt = ktiles(j);
          for (int t = 0; t < nt; t++, kt++) {
            const int s = kt % SL_STAGES, it = kt / SL_STAGES;
            if (it > 0) mbar_wait(smem_u32(&empty_bar[s]), (it - 1) & 1);
            const uint32_t fb = smem_u32(&full_bar[s]);
            mbar_expect_tx(fb, SL_STAGE_BYTES);
            const uint32_t sa = smem_base + s * SL_STAGE_BYTES, sb = sa + SL_TILE_BYTES;
            const int am = p.off + i * SL_BM, k = p.off + j * SL_BM + t * SL_BK;
            if (AMAJ == MAJ_MN) tma_load_3d(sa, &mapT, fb, 0, k, am >> 3);
            else                tma_load_3d(sa, &mapT, fb, 0, am, k >> 3);
            tma_load_3d(sb, &mapV, fb, 0, v0, k >> 3);   // X[k-range, slab] (K-major)
          }
        }
        if (SOLVE) {
          // the diagonal tile of this block row streams through the same ring, 16 columns at a time (A operand only),
          // in the order the substitution consumes it (ascending for lower, descending for upper)
          const int nt = ktiles(i);
          for (int tt = 0; tt < nt; tt++, kt++) {
            const int t = LOWER ? tt : nt - 1 - tt;
            const int s = kt % SL_STAGES, it = kt / SL_STAGES;
            if (it > 0) mbar_wait(smem_u32(&empty_bar[s]), (it - 1) & 1);
            const uint32_t fb = smem_u32(&full_bar[s]);
            mbar_expect_tx(fb, SL_TILE_BYTES);
            const uint32_t sa = smem_base + s * SL_STAGE_BYTES;
            const int am = p.off + i * SL_BM, k = am + t * SL_BK;
            if (AMAJ == MAJ_MN) tma_load_3d(sa, &mapT, fb, 0, k, am >> 3);
            else                tma_load_3d(sa, &mapT, fb, 0, am, k >> 3);
          }
        }
      }
    }
    return;
  }

  // ===== consumers: warp w owns vectors [16w, 16w+16) of the slab, all 128 rows =====
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  const uint32_t g = lane >> 2, q = lane & 3;
  uint32_t aoff[4], boff[4];
  uint32_t kloc[4];  // k index (0..15) inside the tile handled by this lane at each step
#pragma unroll
  for (int st = 0; st < 4; st++) {
    const uint32_t s1 = st & 1, s2 = st >> 1;
    const uint32_t ki = (q & 1) + 4 * (q >> 1) + 2 * s1;
    const uint32_t k = ki + 8 * s2;
    kloc[st] = k;
    if (AMAJ == MAJ_MN) aoff[st] = sw64(k, g);                               // + i16*1024
    else                aoff[st] = s2 * (SL_BM * 64u) + sw64(g, ki);         // + i16*512
    boff[st] = SL_TILE_BYTES + s2 * (SL_W * 64u) + sw64(warp * 16 + g, ki);  // + jn*512
  }
  constexpr uint32_t ASTR = (AMAJ == MAJ_MN) ? 1024u : 512u;
  double* scratch = reinterpret_cast<double*>(smem_gen + SL_STAGES * SL_STAGE_BYTES) + warp * 128;

  double acc[16][2][2];
  int kt = 0;
  for (int r = 0; r < nb; r++) {
    const int i = ASC ? r : nb - 1 - r;
#pragma unroll
    for (int a = 0; a < 16; a++)
#pragma unroll
      for (int b = 0; b < 2; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

    const int cnt = kcount(i);
    for (int jj = 0; jj < cnt; jj++) {
      const int j = kblock(i, jj);
      const bool diag = !SOLVE && (j == i);
      const int nt = ktiles(j);
      for (int t = 0; t < nt; t++, kt++) {
        const int s = kt % SL_STAGES;
        mbar_wait(smem_u32(&full_bar[s]), (kt / SL_STAGES) & 1);
        const uint32_t sbase = smem_base + s * SL_STAGE_BYTES;
#pragma unroll
        for (int st = 0; st < 4; st++) {
          double b[2];
#pragma unroll
          for (int y = 0; y < 2; y++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(b[y]) : "r"(sbase + boff[st] + y * 512u));
          const int kk = t * SL_BK + (int)kloc[st];
#pragma unroll
          for (int h = 0; h < 2; h++) {  // two halves of 8 row-blocks keep the live A fragments at 8
            double a[8];
#pragma unroll
            for (int x = 0; x < 8; x++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a[x]) : "r"(sbase + aoff[st] + (h * 8 + x) * ASTR));
            if (diag) {  // triangular K block of a multiply: keep only the `uplo` part of the diagonal tile
#pragma unroll
              for (int x = 0; x < 8; x++) {
                const int row = (h * 8 + x) * 8 + (int)g;
                if (LOWER ? (row < kk) : (row > kk)) a[x] = 0.0;
                if (p.unit && row == kk) a[x] = 1.0;
              }
            }
#pragma unroll
            for (int x = 0; x < 8; x++)
#pragma unroll
              for (int y = 0; y < 2; y++) dmma884(acc[h * 8 + x][y][0], acc[h * 8 + x][y][1], a[x], b[y]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty_bar[s]));
      }
    }

    // ===== block-row epilogue =====
    const int rbase = i * SL_BM;                 // block-local row offset inside the diagonal block
    const int vrows = min(SL_BM, p.T - rbase);   // valid rows of this block row
    double* Bblk = p.B + (long long)(p.off + rbase);
    const int col0 = v0 + warp * 16 + 2 * (int)q;

    if (SOLVE) {
      // rhs = beta*B - S
#pragma unroll
      for (int y = 0; y < 2; y++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const int col = col0 + y * 8 + c;
          const double* bp = Bblk + (long long)col * p.ldb;
#pragma unroll
          for (int x = 0; x < 16; x++) {
            const int lr = x * 8 + (int)g;
            double b = 0.0;
            if (col < v_end && lr < vrows) b = bp[lr];
            if (p.beta != 1.0) b = __dmul_rn(p.beta, b);
            acc[x][y][c] = b - acc[x][y][c];
          }
        }
      // in-register triangular solve of the 128x128 diagonal tile: 16-column tiles of it arrive through the ring;
      // each holds two 8-row micro-blocks.  The loops stay rolled (code size); accumulator fragments are picked with
      // warp-uniform predicates over statically indexed registers so that `acc` never leaves the register file.
      const int ndt = ktiles(i);
#pragma unroll 1
      for (int tt = 0; tt < ndt; tt++, kt++) {
        const int tc = LOWER ? tt : ndt - 1 - tt;      // 16-column tile of the diagonal block
        const int s = kt % SL_STAGES;
        mbar_wait(smem_u32(&full_bar[s]), (kt / SL_STAGES) & 1);
        const uint32_t sbase = smem_base + s * SL_STAGE_BYTES;
#pragma unroll 1
        for (int hh = 0; hh < 2; hh++) {
          const int s2 = LOWER ? hh : 1 - hh;          // which 8 columns of the tile
          const int ib = tc * 2 + s2;                  // micro-block = row-block index inside the 128 tile
          const int mr = ib * 8;
          const bool valid = (mr + (int)g) < vrows;
          // row g of the 8x8 diagonal micro-block, from the staged tile (zero beyond the matrix edge)
          double lrow[8];
#pragma unroll
          for (int pp = 0; pp < 8; pp++) {
            const uint32_t o = (AMAJ == MAJ_MN) ? ((uint32_t)ib * 1024u + sw64(8u * s2 + pp, g))
                                                : ((uint32_t)s2 * (SL_BM * 64u) + sw64((uint32_t)ib * 8u + g, pp));
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(lrow[pp]) : "r"(sbase + o));
          }
          double d = 1.0;
#pragma unroll
          for (int pp = 0; pp < 8; pp++) if (pp == (int)g && valid && !p.unit) d = lrow[pp];
          // the reference scales every entry of a row by its diagonal (src/trsm.jl:15-18,24); one reciprocal per row
          // instead of 12 divisions keeps the FP64 pipe for the DMMAs (costs <= 1 ulp per scaled entry)
          const double rd = 1.0 / d;
          double lp[8];
#pragma unroll
          for (int pp = 0; pp < 8; pp++) {
            const bool dep = LOWER ? (pp < (int)g) : (pp > (int)g);
            lp[pp] = (valid && dep) ? lrow[pp] * rd : 0.0;
          }
          double x[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#define NLA_GET(XX) case XX: x[0][0] = acc[XX][0][0]; x[0][1] = acc[XX][0][1]; x[1][0] = acc[XX][1][0]; x[1][1] = acc[XX][1][1]; break;
          switch (ib) {   // warp-uniform: one taken case instead of 16 predicated copies
            NLA_GET(0) NLA_GET(1) NLA_GET(2) NLA_GET(3) NLA_GET(4) NLA_GET(5) NLA_GET(6) NLA_GET(7)
            NLA_GET(8) NLA_GET(9) NLA_GET(10) NLA_GET(11) NLA_GET(12) NLA_GET(13) NLA_GET(14) NLA_GET(15)
          }
#undef NLA_GET
#pragma unroll
          for (int y = 0; y < 2; y++)
#pragma unroll
            for (int c = 0; c < 2; c++) x[y][c] = valid ? x[y][c] * rd : 0.0;
          // warp-shuffle substitution: pivot row pv lives in the lanes with g == pv
#pragma unroll
          for (int pp = 0; pp < 8; pp++) {
            const int pv = LOWER ? pp : 7 - pp;
            const double l = lp[pv];
#pragma unroll
            for (int y = 0; y < 2; y++)
#pragma unroll
              for (int c = 0; c < 2; c++) {
                const double xp = __shfl_sync(0xffffffffu, x[y][c], 4 * pv + (int)q);
                x[y][c] = fma(-l, xp, x[y][c]);
              }
          }
#define NLA_PUT(XX) case XX: acc[XX][0][0] = x[0][0]; acc[XX][0][1] = x[0][1]; acc[XX][1][0] = x[1][0]; acc[XX][1][1] = x[1][1]; break;
          switch (ib) {
            NLA_PUT(0) NLA_PUT(1) NLA_PUT(2) NLA_PUT(3) NLA_PUT(4) NLA_PUT(5) NLA_PUT(6) NLA_PUT(7)
            NLA_PUT(8) NLA_PUT(9) NLA_PUT(10) NLA_PUT(11) NLA_PUT(12) NLA_PUT(13) NLA_PUT(14) NLA_PUT(15)
          }
#undef NLA_PUT

          // X_ib (8 x 16, accumulator layout) -> B fragments (k permuted like the main loop) via the warp's scratch
          __syncwarp();
#pragma unroll
          for (int y = 0; y < 2; y++)
#pragma unroll
            for (int c = 0; c < 2; c++) scratch[g * 16 + y * 8 + 2 * q + c] = x[y][c];
          __syncwarp();
          double bf[2][2];
#pragma unroll
          for (int y = 0; y < 2; y++)
#pragma unroll
            for (int s1 = 0; s1 < 2; s1++) bf[y][s1] = scratch[((q & 1) + 4 * (q >> 1) + 2 * s1) * 16 + y * 8 + g];
          // rows still to be solved: rhs -= Teff[row-block, micro-block] * X_ib, A fragments from the staged tile
          const uint32_t ao0 = s2 ? aoff[2] : aoff[0], ao1 = s2 ? aoff[3] : aoff[1];
#pragma unroll
          for (int xo = 0; xo < 16; xo++) {
            const int xx = LOWER ? xo : 15 - xo;
            const bool after = LOWER ? (xx > ib) : (xx < ib);
            if (after) {
              double a0, a1;
              asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a0) : "r"(sbase + ao0 + xx * ASTR));
              asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a1) : "r"(sbase + ao1 + xx * ASTR));
              a0 = -a0; a1 = -a1;
#pragma unroll
              for (int y = 0; y < 2; y++) {
                dmma884(acc[xx][y][0], acc[xx][y][1], a0, bf[y][0]);
                dmma884(acc[xx][y][0], acc[xx][y][1], a1, bf[y][1]);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty_bar[s]));
      }
    }

    // write the block row back in place
#pragma unroll
    for (int y = 0; y < 2; y++)
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const int col = col0 + y * 8 + c;
        if (col < v_end) {
          double* bp = Bblk + (long long)col * p.ldb;
#pragma unroll
          for (int x = 0; x < 16; x++) {
            const int lr = x * 8 + (int)g;
            if (lr < vrows) {
              double v = acc[x][y][c];
              if (!SOLVE && p.post != 1.0) v = __dmul_rn(p.post, v);
              bp[lr] = v;
            }
          }
        }
      }
    if (SOLVE) {
      __threadfence();
      fence_proxy_async_all();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&xdone_bar));
    }
  }
}

}  // namespace nla

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

constexpr int SL_BM = 128, SL_BK = 16;
constexpr int SL_TILE_BYTES = SL_BM * SL_BK * 8;                        // one 128 x 16 tile of Teff
constexpr int SL_SCR_PITCH = 9;                                        // doubles per column of a warp's 8 x 16 exchange tile (8 rows + 1 pad)
constexpr int SL_LP_DOUBLES = 16 * 64 + 128;                            // per block row: 16 scaled 8 x 8 micro-blocks + 128 reciprocal diagonals
constexpr int SL_LP_BYTES = 2 * SL_LP_DOUBLES * 8;                      // double-buffered by block-row parity

// Machine mapping: a CTA owns W right-hand-side vectors, one consumer warp per 16 of them, plus a warpgroup that holds the TMA producer
// warp and the helper warp and hands most of its registers to the consumers (setmaxnreg).
//   W = 128: one CTA per SM (8 consumer warps, 4-stage ring).
//   W = 64 : TWO CTAs per SM (4 consumer warps each, 3-stage rings).  The two CTAs of an SM are independent (different vectors) and drift
//            apart, so one runs its latency-bound diagonal phase / write-back while the other keeps the FP64 tensor pipe busy.
template <int W> struct SlabCfg {
  static constexpr int CW = W / 16;
  static constexpr int THREADS = (CW + 4) * 32;
  static constexpr int CONSUMER_REGS = W == 128 ? 232 : 216;   // 12 warps x 168 -> 8 x 232 + 4 x 40;   2 CTAs x (8 warps x 128 -> 4 x 216 + 4 x 40)
  static constexpr int STAGES = W == 128 ? 4 : 3;
  static constexpr int B_BYTES = W * SL_BK * 8;
  static constexpr int STAGE_BYTES = SL_TILE_BYTES + B_BYTES;
  static constexpr int SCRATCH_BYTES = CW * 16 * SL_SCR_PITCH * 8;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + SCRATCH_BYTES + SL_LP_BYTES + 1024;
  static constexpr int MIN_CTAS = W == 128 ? 1 : 2;
};

struct SlabParams {
  int T;                  // order of the diagonal block
  int off;                // its origin (row == column) in Teff coordinates
  int v_base, v_count;    // vectors (columns of B) handled by this launch
  long long m_total;      // vectors of the whole call (host side: picks the CTA width)
  const double* A;        // Teff(r,c) = A[r*t_rs + c*t_cs]
  long long t_rs, t_cs;
  double* B;              // column-major, vectors are columns
  long long ldb;
  double beta, post;
  int unit;               // 1: unit diagonal (the stored diagonal is not used)
  unsigned long long* dbg;   // probes only (option tc_dbg): 4 globaltimer stamps per block row from warp 0 of CTA 0
  // Streaming mode of the row-split kernel (host-buffer pipeline, nla_api.cu host_stream_pipeline): the operands are still ARRIVING from
  // the host while the kernel runs and finished rows LEAVE while it runs.  All null / zero for device-resident calls.
  const int* flag_in;        // device word, bumped (stream-ordered, after the copies) every time another chunk of A and B has landed
  const int* need;           // [block row in processing order] value flag_in must have reached before that block row may start
  const int* chunk_of;       // [block row in processing order] output chunk the block row belongs to
  const int* chunk_total;    // [chunk] warp arrivals that complete the chunk (= its block rows x CTAs x consumer warps)
  int* done_cnt;             // [chunk] arrival counters (zeroed by the host)
  volatile int* flag_out;    // [chunk] host-mapped words: set to 1 when the chunk is final in device memory (the host then downloads it)
};

// streaming mode: spin until *flag >= need (acquire: data copied in before the flag was bumped is visible afterwards)
__device__ __forceinline__ void slab_wait_flag(const int* flag, int need) {
  int v;
  do {
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
  } while (v < need);
}

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int AMAJ, bool LOWER, bool SOLVE, int W>
__global__ void __launch_bounds__(SlabCfg<W>::THREADS, SlabCfg<W>::MIN_CTAS)
slab_f64_kernel(const __grid_constant__ CUtensorMap mapT, const __grid_constant__ CUtensorMap mapV, const SlabParams p) {
  using Cfg = SlabCfg<W>;
  constexpr int SL_W = W, SL_STAGES = Cfg::STAGES, SL_STAGE_BYTES = Cfg::STAGE_BYTES, SL_CONSUMER_WARPS = Cfg::CW, SL_SCRATCH_BYTES = Cfg::SCRATCH_BYTES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[SL_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[SL_STAGES];
  __shared__ __align__(8) uint64_t xdone_bar;
  __shared__ __align__(8) uint64_t lp_full[2];    // scaled micro-blocks of a block row are in shared memory (helper warp -> consumers)
  __shared__ __align__(8) uint64_t lp_empty[2];   // every consumer warp has finished with them
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SL_STAGES; s++) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), SL_CONSUMER_WARPS);
    }
    mbar_init(smem_u32(&xdone_bar), SL_CONSUMER_WARPS);
    for (int b = 0; b < 2; b++) {
      mbar_init(smem_u32(&lp_full[b]), 1);
      mbar_init(smem_u32(&lp_empty[b]), SL_CONSUMER_WARPS);
    }
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
    // ===== producer warpgroup: hands its registers to the consumers; warp 0 of it issues the TMA copies (one lane), warp 1 is the helper =====
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
    if (SOLVE && warp == SL_CONSUMER_WARPS + 1) {
      // ===== helper warp: the 8 x 8 diagonal micro-blocks of every block row in the reference's scaled form (src/trsm.jl:15-18,24):
      // reciprocal diagonal rd_r = 1/d_r and l'_rk = a_rk * rd_r.  They depend on A only, so this otherwise idle warp prepares them one
      // block row AHEAD of the consumers, straight from global memory into a double-buffered shared array: the global loads and the FP64
      // divisions never appear in the consumers' dependency chain. =====
      uint8_t* gen = smem_raw + (smem_base - smem_u32(smem_raw));
      double* lp_base = reinterpret_cast<double*>(gen + SL_STAGES * SL_STAGE_BYTES + SL_SCRATCH_BYTES);
      const int hg = lane >> 2, hq = lane & 3;
      for (int r = 0; r < nb; r++) {
        const int i = ASC ? r : nb - 1 - r;
        if (r >= 2) mbar_wait(smem_u32(&lp_empty[r & 1]), ((r >> 1) - 1) & 1);
        double* lpw = lp_base + (r & 1) * SL_LP_DOUBLES;
        const int vr = min(SL_BM, p.T - i * SL_BM);
        const long long base = (long long)p.off + i * SL_BM;
#pragma unroll 4
        for (int mb = 0; mb < 16; mb++) {
          const int rr = mb * 8 + hg;                        // row inside the block row
          double v[2];
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int cc = mb * 8 + 2 * hq + e;              // column inside the block row
            v[e] = 0.0;
            if (rr < vr && cc < vr) v[e] = p.A[(base + rr) * p.t_rs + (base + cc) * p.t_cs];
          }
          // the diagonal entry of row g sits in lane 4g + (g >> 1), element g & 1
          const double d0 = __shfl_sync(0xffffffffu, v[0], 4 * hg + (hg >> 1));
          const double d1 = __shfl_sync(0xffffffffu, v[1], 4 * hg + (hg >> 1));
          const bool valid = rr < vr;
          const double d = (valid && !p.unit) ? ((hg & 1) ? d1 : d0) : 1.0;
          // one reciprocal per row instead of a division per entry (costs <= 1 ulp per scaled entry)
          const double rd = 1.0 / d;
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int c8 = 2 * hq + e;
            const bool dep = LOWER ? (c8 < hg) : (c8 > hg);
            lpw[mb * 64 + hg * 8 + c8] = (valid && dep) ? v[e] * rd : 0.0;
          }
          if (hq == 0) lpw[16 * 64 + rr] = valid ? rd : 0.0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&lp_full[r & 1]));
      }
    }
    return;
  }

  // ===== consumers: warp w owns vectors [16w, 16w+16) of the slab, all 128 rows =====
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::CONSUMER_REGS));
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
  double* scratch = reinterpret_cast<double*>(smem_gen + SL_STAGES * SL_STAGE_BYTES) + warp * (16 * SL_SCR_PITCH);
  double* lp_all = reinterpret_cast<double*>(smem_gen + SL_STAGES * SL_STAGE_BYTES + SL_SCRATCH_BYTES);

#define NLA_SLAB_STAMP(slot) do { if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[(slot)] = global_timer_ns(); } while (0)
  double acc[16][2][2];
  int kt = 0;
  for (int r = 0; r < nb; r++) {
    const int i = ASC ? r : nb - 1 - r;
    NLA_SLAB_STAMP(4 * r);
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
    NLA_SLAB_STAMP(4 * r + 1);
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
      const double* lpr = lp_all + (r & 1) * SL_LP_DOUBLES;
      mbar_wait(smem_u32(&lp_full[r & 1]), (r >> 1) & 1);   // the helper warp has this block row's scaled micro-blocks in shared memory
      // In-register triangular solve of the 128 x 128 diagonal tile.  16-column tiles of it arrive through the ring (ascending for a
      // lower, descending for an upper triangle); each holds two 8-row micro-blocks.  The 16 micro-block steps are FULLY UNROLLED: the
      // micro-block index is a compile-time constant, so its accumulator fragments are named registers and the trailing update is
      // straight-line code.  (A rolled loop has to pick fragments with predicates, and predicated-off DMMAs still occupy the FP64
      // tensor pipe: measured, every micro-block then pays for all 15 row-blocks.)
#pragma unroll
      for (int tcu = 0; tcu < SL_BM / SL_BK; tcu++) {
        constexpr int NT = SL_BM / SL_BK;
        const int tc = LOWER ? tcu : NT - 1 - tcu;     // 16-column tile of the diagonal block (compile-time)
        if (tc < ndt) {                                 // ragged last block row: tiles beyond the matrix edge do not exist
          const int s = kt % SL_STAGES;
          mbar_wait(smem_u32(&full_bar[s]), (kt / SL_STAGES) & 1);
          const uint32_t sbase = smem_base + s * SL_STAGE_BYTES;
#pragma unroll
          for (int hh = 0; hh < 2; hh++) {
            const int s2 = LOWER ? hh : 1 - hh;        // which 8 columns of the tile
            const int ib = tc * 2 + s2;                // micro-block = row-block index inside the 128 tile (compile-time)
            // The 8 x 16 right-hand-side micro-block goes through the warp's exchange tile (column-major, pitch 9): lane c < 16 then
            // owns all 8 rows of vector c and runs the substitution of the reference's leaf (src/trsm.jl:15-27: x_r = b_r/d_r -
            // sum l'_rk x_k, pivots in order) in its own registers -- 28 FMAs, a dependent chain of 7 -- instead of 8 shuffle rounds.
            __syncwarp();
#pragma unroll
            for (int y = 0; y < 2; y++)
#pragma unroll
              for (int c = 0; c < 2; c++) scratch[(y * 8 + 2 * (int)q + c) * SL_SCR_PITCH + (int)g] = acc[ib][y][c];
            __syncwarp();
            if (lane < 16) {
              const double* lm = lpr + ib * 64;
              const double* rv = lpr + 16 * 64 + ib * 8;
              double* col = scratch + lane * SL_SCR_PITCH;
              double b[8];
#pragma unroll
              for (int rr = 0; rr < 8; rr++) b[rr] = col[rr] * rv[rr];
#pragma unroll
              for (int pp = 0; pp < 8; pp++) {
                const int pv = LOWER ? pp : 7 - pp;    // pivot
#pragma unroll
                for (int rr = 0; rr < 8; rr++)
                  if (LOWER ? (rr > pv) : (rr < pv)) b[rr] = fma(-lm[rr * 8 + pv], b[pv], b[rr]);
              }
#pragma unroll
              for (int rr = 0; rr < 8; rr++) col[rr] = b[rr];
            }
            __syncwarp();
#pragma unroll
            for (int y = 0; y < 2; y++)
#pragma unroll
              for (int c = 0; c < 2; c++) acc[ib][y][c] = scratch[(y * 8 + 2 * (int)q + c) * SL_SCR_PITCH + (int)g];
            // X_ib as B fragments (k permuted like the main loop)
            double bf[2][2];
#pragma unroll
            for (int y = 0; y < 2; y++)
#pragma unroll
              for (int s1 = 0; s1 < 2; s1++) bf[y][s1] = scratch[(y * 8 + (int)g) * SL_SCR_PITCH + (q & 1) + 4 * (q >> 1) + 2 * s1];
            // rows still to be solved: rhs -= Teff[row-block, micro-block] * X_ib, A fragments from the staged tile, four row-blocks
            // at a time (loads first, then their 8 independent DMMAs); the second k-half follows a whole pass later, so the two DMMAs
            // that hit the same accumulator are never back to back
            const int NTR = LOWER ? 15 - ib : ib;      // row-blocks still to be updated (a constant once the loops are unrolled)
#pragma unroll
            for (int s1 = 0; s1 < 2; s1++) {
              const uint32_t ao = s2 ? aoff[2 + s1] : aoff[s1];
#pragma unroll
              for (int t0 = 0; t0 < 16; t0 += 4) {
                double av[4];
#pragma unroll
                for (int u = 0; u < 4; u++)
                  if (t0 + u < NTR) {
                    const int xx = LOWER ? ib + 1 + t0 + u : ib - 1 - t0 - u;
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(av[u]) : "r"(sbase + ao + xx * ASTR));
                    av[u] = -av[u];
                  }
#pragma unroll
                for (int u = 0; u < 4; u++)
                  if (t0 + u < NTR) {
                    const int xx = LOWER ? ib + 1 + t0 + u : ib - 1 - t0 - u;
                    dmma884(acc[xx][0][0], acc[xx][0][1], av[u], bf[0][s1]);
                    dmma884(acc[xx][1][0], acc[xx][1][1], av[u], bf[1][s1]);
                  }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&empty_bar[s]));
          kt++;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&lp_empty[r & 1]));
    }

    // write the block row back in place
    NLA_SLAB_STAMP(4 * r + 2);
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
    NLA_SLAB_STAMP(4 * r + 3);
  }
#undef NLA_SLAB_STAMP
}

}  // namespace nla

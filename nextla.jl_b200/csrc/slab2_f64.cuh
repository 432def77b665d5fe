// slab2_f64.cuh -- fused FP64 macro-leaf, ROW-SPLIT mapping: the successor of slab_f64.cuh for the machine-filling case.
//
// Same job as slab_f64_kernel (one CTA solves / multiplies a whole T x T diagonal block against its own slab of right-hand-side
// vectors, left-looking over 128-row block rows, no inter-CTA communication; replaces the reference's leaves src/trsm.jl:5-126,
// src/trmm.jl:43-312 and the small-K GEMM levels above them, src/rectrxm.jl:159-197), but the 8 consumer warps split the 128 ROWS of a
// block row between them and every warp works on ALL the vectors of the CTA.  The column-split kernel needs a multiple of 4 column
// blocks per SM sub-partition to keep the four FP64 tensor pipes equally loaded, i.e. 128 vectors per CTA -- and 16384 vectors are 128
// CTAs on 148 SMs (86 % of the machine).  With rows split, the CTA width is any multiple of 8:
//     W = 112 : 16384 vectors = 147 CTAs (4 x 4096-vector stream slabs = 4 x 37 = 148 CTAs),   W = 56 : 8192 vectors = 147 CTAs.
// Warp w owns the 8-row blocks w and 15 - w of the block row ("snake": every warp then has the same amount of trailing-update work in
// the triangular phase), accumulators acc[2][NB][2] (NB = W / 8 column blocks).
//
// Main loop: as in gemm_f64.cuh / slab_f64.cuh (TMA -> 4-stage mbarrier ring -> DMMA.8x8x4, conflict-free k permutation); per k-quad a
// warp loads 2 A fragments and NB B fragments for 2 NB DMMAs.
// Triangular phase of a block row (solve): 16 micro-block steps, fully unrolled (compile-time micro-block index, see slab_f64.cuh for
// why).  Step ib: (A) the owner warp writes its 8 x W right-hand-side micro-block into a column-major exchange tile; named barrier;
// (B) thread c < W runs the substitution of the reference's leaf (src/trsm.jl:15-27, scaled form, pivots in order) for vector c in its
// own registers, with the scaled micro-block prepared ahead of time by the helper warp; named barrier; (C) the owner reads its solved
// rows back, every warp whose row blocks come later loads X_ib as DMMA B fragments from the exchange tile and updates its rows.  The
// exchange tile is double-buffered by micro-block parity, so two barriers per step are enough.
#pragma once
#include "slab_f64.cuh"

namespace nla {

template <int NB> struct Slab2Cfg {
  static constexpr int W = 8 * NB;
  static constexpr int CW = 8;                                   // consumer warps
  static constexpr int THREADS = (CW + 4) * 32;                  // + producer warpgroup (TMA warp, helper warp, two idle warps)
  static constexpr int STAGES = 4;
  static constexpr int B_BYTES = W * SL_BK * 8;
  static constexpr int STAGE_BYTES = SL_TILE_BYTES + B_BYTES;    // a multiple of 1024 for every NB (16384 + NB * 1024)
  static constexpr int XCH_DOUBLES = W * SL_SCR_PITCH;           // one exchange tile
  static constexpr int XCH_BYTES = 2 * XCH_DOUBLES * 8;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + XCH_BYTES + SL_LP_BYTES + 1024;
};

template <int AMAJ, bool LOWER, bool SOLVE, int NB>
__global__ void __launch_bounds__(Slab2Cfg<NB>::THREADS, 1)
slab2_f64_kernel(const __grid_constant__ CUtensorMap mapT, const __grid_constant__ CUtensorMap mapV, const SlabParams p) {
  using Cfg = Slab2Cfg<NB>;
  constexpr int W = Cfg::W, STAGES = Cfg::STAGES, STAGE_BYTES = Cfg::STAGE_BYTES, CW = Cfg::CW;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t xdone_bar;
  __shared__ __align__(8) uint64_t lp_full[2];
  __shared__ __align__(8) uint64_t lp_empty[2];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), CW);
    }
    mbar_init(smem_u32(&xdone_bar), CW);
    for (int b = 0; b < 2; b++) {
      mbar_init(smem_u32(&lp_full[b]), 1);
      mbar_init(smem_u32(&lp_empty[b]), CW);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int nb = (p.T + SL_BM - 1) / SL_BM;
  constexpr bool ASC = (SOLVE == LOWER);
  const int v0 = p.v_base + blockIdx.x * W;
  const int v_end = p.v_base + p.v_count;
  auto kcount = [&](int i) { return SOLVE ? (LOWER ? i : nb - 1 - i) : (LOWER ? i + 1 : nb - i); };
  auto kblock = [&](int i, int jj) { return SOLVE ? (LOWER ? jj : nb - 1 - jj) : (LOWER ? jj : i + jj); };
  auto ktiles = [&](int j) { return (min(SL_BM, p.T - j * SL_BM) + SL_BK - 1) / SL_BK; };

  if (warp >= CW) {
    // ===== producer warpgroup: hands its registers to the consumers; warp 0 of it issues the TMA copies, warp 1 is the helper =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == CW && lane == 0) {
      tma_prefetch_desc(&mapT);
      tma_prefetch_desc(&mapV);
      int kt = 0;
      for (int r = 0; r < nb; r++) {
        const int i = ASC ? r : nb - 1 - r;
        if (p.flag_in) slab_wait_flag(p.flag_in, p.need[r]);   // streaming mode: this block row's rows of A have landed
        const int cnt = kcount(i);
        for (int jj = 0; jj < cnt; jj++) {
          const int j = kblock(i, jj);
          if (SOLVE && jj == cnt - 1) mbar_wait(smem_u32(&xdone_bar), (r - 1) & 1);  // rows written by the previous block row
          const int nt = ktiles(j);
          for (int t = 0; t < nt; t++, kt++) {
            const int s = kt % STAGES, it = kt / STAGES;
            if (it > 0) mbar_wait(smem_u32(&empty_bar[s]), (it - 1) & 1);
            const uint32_t fb = smem_u32(&full_bar[s]);
            mbar_expect_tx(fb, STAGE_BYTES);
            const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + SL_TILE_BYTES;
            const int am = p.off + i * SL_BM, k = p.off + j * SL_BM + t * SL_BK;
            if (AMAJ == MAJ_MN) tma_load_3d(sa, &mapT, fb, 0, k, am >> 3);
            else                tma_load_3d(sa, &mapT, fb, 0, am, k >> 3);
            tma_load_3d(sb, &mapV, fb, 0, v0, k >> 3);   // X[k-range, slab] (K-major), box {8, W, 2}
          }
        }
        if (SOLVE) {   // the diagonal tile of this block row, 16 columns at a time (A operand only), in consumption order
          const int nt = ktiles(i);
          for (int tt = 0; tt < nt; tt++, kt++) {
            const int t = LOWER ? tt : nt - 1 - tt;
            const int s = kt % STAGES, it = kt / STAGES;
            if (it > 0) mbar_wait(smem_u32(&empty_bar[s]), (it - 1) & 1);
            const uint32_t fb = smem_u32(&full_bar[s]);
            mbar_expect_tx(fb, SL_TILE_BYTES);
            const uint32_t sa = smem_base + s * STAGE_BYTES;
            const int am = p.off + i * SL_BM, k = am + t * SL_BK;
            if (AMAJ == MAJ_MN) tma_load_3d(sa, &mapT, fb, 0, k, am >> 3);
            else                tma_load_3d(sa, &mapT, fb, 0, am, k >> 3);
          }
        }
      }
    }
    if (SOLVE && warp == CW + 1) {
      // ===== helper warp: scaled 8 x 8 diagonal micro-blocks (l'_rk = a_rk / d_r, rd_r = 1 / d_r; src/trsm.jl:15-18,24) one block row
      // ahead of the consumers, from global memory into a double-buffered shared array =====
      double* lp_base = reinterpret_cast<double*>(smem_gen + STAGES * STAGE_BYTES + Cfg::XCH_BYTES);
      const int hg = lane >> 2, hq = lane & 3;
      for (int r = 0; r < nb; r++) {
        const int i = ASC ? r : nb - 1 - r;
        if (r >= 2) mbar_wait(smem_u32(&lp_empty[r & 1]), ((r >> 1) - 1) & 1);
        if (p.flag_in) { if (lane == 0) slab_wait_flag(p.flag_in, p.need[r]); __syncwarp(); }
        double* lpw = lp_base + (r & 1) * SL_LP_DOUBLES;
        const int vr = min(SL_BM, p.T - i * SL_BM);
        const long long base = (long long)p.off + i * SL_BM;
#pragma unroll 4
        for (int mb = 0; mb < 16; mb++) {
          const int rr = mb * 8 + hg;
          double v[2];
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int cc = mb * 8 + 2 * hq + e;
            v[e] = 0.0;
            if (rr < vr && cc < vr) v[e] = p.A[(base + rr) * p.t_rs + (base + cc) * p.t_cs];
          }
          const double d0 = __shfl_sync(0xffffffffu, v[0], 4 * hg + (hg >> 1));
          const double d1 = __shfl_sync(0xffffffffu, v[1], 4 * hg + (hg >> 1));
          const bool valid = rr < vr;
          const double d = (valid && !p.unit) ? ((hg & 1) ? d1 : d0) : 1.0;
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

  // ===== consumers: warp w owns the 8-row blocks w and 15 - w of every block row, all W vectors =====
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  const uint32_t g = lane >> 2, q = lane & 3;
  const int mbk[2] = {warp, 15 - warp};                     // this warp's row blocks
  uint32_t aoff[4], boff[4];
  uint32_t kloc[4];
#pragma unroll
  for (int st = 0; st < 4; st++) {
    const uint32_t s1 = st & 1, s2 = st >> 1;
    const uint32_t ki = (q & 1) + 4 * (q >> 1) + 2 * s1;
    const uint32_t k = ki + 8 * s2;
    kloc[st] = k;
    if (AMAJ == MAJ_MN) aoff[st] = sw64(k, g);                               // + rowblock * 1024
    else                aoff[st] = s2 * (SL_BM * 64u) + sw64(g, ki);         // + rowblock * 512
    boff[st] = SL_TILE_BYTES + s2 * (W * 64u) + sw64(g, ki);                 // + jn * 512
  }
  constexpr uint32_t ASTR = (AMAJ == MAJ_MN) ? 1024u : 512u;
  double* xch = reinterpret_cast<double*>(smem_gen + STAGES * STAGE_BYTES);
  const double* lp_all = reinterpret_cast<const double*>(smem_gen + STAGES * STAGE_BYTES + Cfg::XCH_BYTES);
  const int ctid = threadIdx.x;                              // 0 .. 255 among the consumers

#define NLA_SLAB2_STAMP(slot) do { if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[(slot)] = global_timer_ns(); } while (0)
  double acc[2][NB][2];
  int kt = 0;
  for (int r = 0; r < nb; r++) {
    const int i = ASC ? r : nb - 1 - r;
    NLA_SLAB2_STAMP(4 * r);
#pragma unroll
    for (int x = 0; x < 2; x++)
#pragma unroll
      for (int jn = 0; jn < NB; jn++) acc[x][jn][0] = acc[x][jn][1] = 0.0;

    const int cnt = kcount(i);
    for (int jj = 0; jj < cnt; jj++) {
      const int j = kblock(i, jj);
      const bool diag = !SOLVE && (j == i);
      const int nt = ktiles(j);
      for (int t = 0; t < nt; t++, kt++) {
        const int s = kt % STAGES;
        mbar_wait(smem_u32(&full_bar[s]), (kt / STAGES) & 1);
        const uint32_t sbase = smem_base + s * STAGE_BYTES;
#pragma unroll
        for (int st = 0; st < 4; st++) {
          double a[2];
#pragma unroll
          for (int x = 0; x < 2; x++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a[x]) : "r"(sbase + aoff[st] + (uint32_t)mbk[x] * ASTR));
          if (diag) {  // triangular K block of a multiply: keep only the `uplo` part of the diagonal tile
            const int kk = t * SL_BK + (int)kloc[st];
#pragma unroll
            for (int x = 0; x < 2; x++) {
              const int row = mbk[x] * 8 + (int)g;
              if (LOWER ? (row < kk) : (row > kk)) a[x] = 0.0;
              if (p.unit && row == kk) a[x] = 1.0;
            }
          }
          double b[NB];
#pragma unroll
          for (int jn = 0; jn < NB; jn++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(b[jn]) : "r"(sbase + boff[st] + jn * 512u));
#pragma unroll
          for (int jn = 0; jn < NB; jn++) {
            dmma884(acc[0][jn][0], acc[0][jn][1], a[0], b[jn]);
            dmma884(acc[1][jn][0], acc[1][jn][1], a[1], b[jn]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty_bar[s]));
      }
    }

    // ===== block-row epilogue =====
    NLA_SLAB2_STAMP(4 * r + 1);
    const int rbase = i * SL_BM;
    const int vrows = min(SL_BM, p.T - rbase);
    double* Bblk = p.B + (long long)(p.off + rbase);

    if (SOLVE) {
      if (p.flag_in) { if (lane == 0) slab_wait_flag(p.flag_in, p.need[r]); __syncwarp(); }   // streaming mode: this block row of B has landed
      // rhs = beta*B - S
#pragma unroll
      for (int jn = 0; jn < NB; jn++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const int col = v0 + jn * 8 + 2 * (int)q + c;
          const double* bp = Bblk + (long long)col * p.ldb;
#pragma unroll
          for (int x = 0; x < 2; x++) {
            const int lr = mbk[x] * 8 + (int)g;
            double b = 0.0;
            if (col < v_end && lr < vrows) b = bp[lr];
            if (p.beta != 1.0) b = __dmul_rn(p.beta, b);
            acc[x][jn][c] = b - acc[x][jn][c];
          }
        }
      const int ndt = ktiles(i);
      const double* lpr = lp_all + (r & 1) * SL_LP_DOUBLES;
      mbar_wait(smem_u32(&lp_full[r & 1]), (r >> 1) & 1);
#pragma unroll
      for (int tcu = 0; tcu < SL_BM / SL_BK; tcu++) {
        constexpr int NT = SL_BM / SL_BK;
        const int tc = LOWER ? tcu : NT - 1 - tcu;     // 16-column tile of the diagonal block (compile-time)
        if (tc < ndt) {                                 // ragged last block row: tiles beyond the matrix edge do not exist
          const int s = kt % STAGES;
          mbar_wait(smem_u32(&full_bar[s]), (kt / STAGES) & 1);
          const uint32_t sbase = smem_base + s * STAGE_BYTES;
#pragma unroll
          for (int hh = 0; hh < 2; hh++) {
            const int s2 = LOWER ? hh : 1 - hh;
            const int ib = tc * 2 + s2;                // micro-block = row block inside the 128 tile (compile-time)
            const int ow = ib < 8 ? ib : 15 - ib;      // the warp that owns it ...
            const int ox = ib < 8 ? 0 : 1;             // ... as its row block ox
            double* E = xch + (ib & 1) * Cfg::XCH_DOUBLES;
            // (A) owner: right-hand-side micro-block -> exchange tile, column-major with pitch 9
            if (warp == ow) {
#pragma unroll
              for (int jn = 0; jn < NB; jn++)
#pragma unroll
                for (int c = 0; c < 2; c++) E[(jn * 8 + 2 * (int)q + c) * SL_SCR_PITCH + (int)g] = acc[ox][jn][c];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // (B) one thread per vector: substitution in registers (scaled form, pivots in order)
            if (ctid < W) {
              const double* lm = lpr + ib * 64;
              const double* rv = lpr + 16 * 64 + ib * 8;
              double* col = E + ctid * SL_SCR_PITCH;
              double b[8];
#pragma unroll
              for (int rr = 0; rr < 8; rr++) b[rr] = col[rr] * rv[rr];
#pragma unroll
              for (int pp = 0; pp < 8; pp++) {
                const int pv = LOWER ? pp : 7 - pp;
#pragma unroll
                for (int rr = 0; rr < 8; rr++)
                  if (LOWER ? (rr > pv) : (rr < pv)) b[rr] = fma(-lm[rr * 8 + pv], b[pv], b[rr]);
              }
#pragma unroll
              for (int rr = 0; rr < 8; rr++) col[rr] = b[rr];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // (C) owner: solved rows back into its accumulators
            if (warp == ow) {
#pragma unroll
              for (int jn = 0; jn < NB; jn++)
#pragma unroll
                for (int c = 0; c < 2; c++) acc[ox][jn][c] = E[(jn * 8 + 2 * (int)q + c) * SL_SCR_PITCH + (int)g];
            }
            // (C) every warp: rows still to be solved -= Teff[rows, micro-block] * X_ib (A fragments from the staged diagonal tile)
            const bool upd0 = LOWER ? (mbk[0] > ib) : (mbk[0] < ib);
            const bool upd1 = LOWER ? (mbk[1] > ib) : (mbk[1] < ib);
            if (upd0 || upd1) {   // warp-uniform
#pragma unroll
              for (int s1 = 0; s1 < 2; s1++) {
                const uint32_t ao = s2 ? aoff[2 + s1] : aoff[s1];
                double bf[NB];
#pragma unroll
                for (int jn = 0; jn < NB; jn++) bf[jn] = E[(jn * 8 + (int)g) * SL_SCR_PITCH + (q & 1) + 4 * (q >> 1) + 2 * s1];
                if (upd0) {
                  double a;
                  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(sbase + ao + (uint32_t)mbk[0] * ASTR));
                  a = -a;
#pragma unroll
                  for (int jn = 0; jn < NB; jn++) dmma884(acc[0][jn][0], acc[0][jn][1], a, bf[jn]);
                }
                if (upd1) {
                  double a;
                  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(sbase + ao + (uint32_t)mbk[1] * ASTR));
                  a = -a;
#pragma unroll
                  for (int jn = 0; jn < NB; jn++) dmma884(acc[1][jn][0], acc[1][jn][1], a, bf[jn]);
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
    NLA_SLAB2_STAMP(4 * r + 2);
#pragma unroll
    for (int jn = 0; jn < NB; jn++)
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const int col = v0 + jn * 8 + 2 * (int)q + c;
        if (col < v_end) {
          double* bp = Bblk + (long long)col * p.ldb;
#pragma unroll
          for (int x = 0; x < 2; x++) {
            const int lr = mbk[x] * 8 + (int)g;
            if (lr < vrows) {
              double v = acc[x][jn][c];
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
      if (p.flag_out && lane == 0) {
        // streaming mode: the warp whose arrival completes an output chunk (all its block rows, every CTA, every consumer warp -- each
        // of them fenced its stores above) tells the host that the chunk may be downloaded
        const int c = p.chunk_of[r];
        if (atomicAdd(p.done_cnt + c, 1) == p.chunk_total[c] - 1) {
          __threadfence_system();
          p.flag_out[c] = 1;
        }
      }
    }
    NLA_SLAB2_STAMP(4 * r + 3);
  }
#undef NLA_SLAB2_STAMP
}

}  // namespace nla

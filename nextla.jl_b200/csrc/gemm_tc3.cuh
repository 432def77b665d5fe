// gemm_tc3.cuh -- PERSISTENT CTA-pair kernel for the Float16 updates and block-inverse leaves (tcgen05.mma.cta_group::2).
//
// Same operation, operands, TMA maps and epilogue arithmetic as gemm_tc.cuh / gemm_tc2.cuh (C <- post*(beta*C + sgn*opA(A)*opB(B)),
// src/matmul.jl:5-81 in the reference).  What changes is the life cycle of a CTA.  ncu on the one-tile-per-CTA kernels for the
// mid-size launches of a solve (profiles/r01_ncu_f16_mid_onetile_summary.csv, r01_ncu_f16_leaf_onetile_summary.csv: M = K = 1024 update 59 us, block-inverse leaf 44 us)
// shows tensor pipe 21-27 % active with DRAM at 10-18 % and L2 at 17-28 %: nothing is saturated, the time goes into per-tile
// fixed costs (barrier/TMEM set-up, an empty operand ring at the start of every tile, a latency-bound drain of 4 warps while the
// tensor core idles).  Here one pair of CTAs per TPC stays resident and walks a static list of 256 x 256 tiles:
//   * the operand ring (6 stages x 32 KB per CTA) never drains: the TMA producer runs ahead into the next tile's K range;
//   * TWO accumulator tiles in tensor memory (2 x 256 columns = all 512): the MMA thread starts tile j+1 as soon as the ring
//     delivers, while 8 drain warps (two per TMEM lane quarter, 128 columns each) write tile j back;
//   * the drain software-pipelines its reads of the old C (the loads of the next 32-column slice are in flight while the
//     current one is converted), and pulls the next tile's C into L2 ahead of time.
// Per-tile K windows 4 / 5 (block-inverse leaves: the pair takes the union of its two 128-row windows -- the extra 128 x 128
// block of the inverse holds explicit zeros) and the second destination `dup` are supported; the batched-TRMM windows 1-3 and
// Float32 (3xTF32 needs the splitter warps and K chunks) stay on the one-tile kernels.
#pragma once
#include "gemm_tc2.cuh"

namespace nla {

struct Tc3Shape {
  static constexpr int BK = 64, UK = 16;
  static constexpr int A_BYTES = TC_BM * BK * 2;      // own 128 rows of A
  static constexpr int B_BYTES = 128 * BK * 2;        // own 128 of the pair's 256 columns of B
  static constexpr int STAGE = A_BYTES + B_BYTES;     // 32 KB
  static constexpr int STAGES = 6;
  static constexpr int DRAIN_WARPS = 8;
  static constexpr int STG_BYTES = DRAIN_WARPS * 4096;   // one 32 x 32 FP32 staging tile per drain warp
  static constexpr int SMEM = STAGES * STAGE + STG_BYTES + 1024;
  static constexpr int THREADS = 64 + 32 * DRAIN_WARPS;
  static constexpr int BN = 256;
  static constexpr int TMEM_COLS = 512;
};

// static tile list of a cluster: pair-tile i -> (pm, tn, kbeg, klen)
struct Tc3Tile { int pm, tn, kbeg, klen; };
__device__ __forceinline__ Tc3Tile tc3_tile(const GemmTcParams& p, int i, int pairs_m) {
  constexpr int GROUP_P = TC_GROUP_M / 2;
  const int per_group = GROUP_P * p.tiles_n;
  const int grp = i / per_group;
  const int first_p = grp * GROUP_P;
  const int gsz = min(GROUP_P, pairs_m - first_p);
  const int rem = i - grp * per_group;
  Tc3Tile t;
  t.pm = first_p + rem % gsz;
  t.tn = rem / gsz;
  t.kbeg = 0; t.klen = p.K;
  if (p.win_mode == 4 || p.win_mode == 5) {
    // triangular windows: the list is sorted by K extent, longest first (tc3_index deals it out in snake order, so every
    // cluster gets the same amount of work to within one tile): w = index along the windowed dimension, o = along the other
    const int no = p.win_on_n ? pairs_m : p.tiles_n, nw = p.win_on_n ? p.tiles_n : pairs_m;
    const int o = i % no;
    int w = i / no;
    if (p.win_mode == 4) w = nw - 1 - w;   // lower: the last block row/column has the longest window
    t.pm = p.win_on_n ? o : w;
    t.tn = p.win_on_n ? w : o;
    if (p.win_mode == 4) {          // lower triangle with its diagonal: k in [0, 256 (w+1))
      t.klen = min(p.K, 256 * (w + 1));
    } else {                        // upper: k in [256 w, K)
      t.kbeg = 256 * w;
      t.klen = p.K - t.kbeg;
    }
  }
  return t;
}
// r-th tile of cluster `cid` (or >= ntiles: none): rounds alternate direction so that sorted lists are dealt out evenly
__device__ __forceinline__ int tc3_index(int r, int cid, int ncl) { return r * ncl + ((r & 1) ? ncl - 1 - cid : cid); }

// Edge tiles of the persistent kernel: one 32 x 32 slice of a warp (rows row0.., columns colg.. of C, `nc` valid columns) from the
// FP32 staging tile stg[column][row]; element-wise, bounds-checked, deliberately not unrolled and not inlined.
__device__ __noinline__ void tc3_store_edge(const GemmTcParams& p, const float* stg, int lane, int row0, int colg, int nc) {
  __half* cbase = reinterpret_cast<__half*>(p.C);
  const int row = row0 + lane;
  if (row >= p.M) return;
#pragma unroll 1
  for (int c = 0; c < nc; c++) {
    __half* dst = cbase + row + (long long)(colg + c) * p.ldc;
    float v = p.overwrite ? 0.f : __half2float(*dst);
    if (p.beta != 1.0f) v = __half2float(__float2half_rn(p.beta * v));
    v += p.sgn * stg[c * 32 + lane];
    if (p.post != 1.0f) v = p.post * __half2float(__float2half_rn(v));
    const __half o = __float2half_rn(v);
    *dst = o;
    const int dr = row - p.dup_r0, dc = colg + c - p.dup_c0;
    if (p.dup && dr >= 0 && dr < p.dup_rn && dc >= 0 && dc < p.dup_cn) reinterpret_cast<__half*>(p.dup)[dr + (long long)dc * p.dup_ld] = o;
  }
}

template <int AMAJ, int BMAJ>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Tc3Shape::THREADS, 1)
gemm_tc3_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const GemmTcParams p) {
  using T = __half;
  using Shp = Tc3Shape;
  constexpr int S = Shp::STAGES, BK = Shp::BK, UK = Shp::UK, BN = Shp::BN;
  constexpr int ES = 2, ATOM = 128 / ES;
  constexpr int A_BYTES = Shp::A_BYTES, STAGE = Shp::STAGE;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[S];      // leader: the bytes of BOTH CTAs have landed
  __shared__ __align__(8) uint64_t empty_bar[S];     // local: the pair's MMAs that read this stage have completed (multicast commit)
  __shared__ __align__(8) uint64_t dfull_bar[2];     // local: a tile has been accumulated into this TMEM buffer (multicast commit)
  __shared__ __align__(8) uint64_t dfree_bar[2];     // leader: both CTAs have drained this TMEM buffer (one arrival per drain warp)
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pairs_m = (p.tiles_m + 1) >> 1;
  int ntiles = pairs_m * p.tiles_n;
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // probes only (option tc_dbg): 64 globaltimer stamps per CTA -- 0 entry, 1 prologue done, per tile j < 14: 4+4j first operand
  // stage landed, 5+4j MMAs issued, 6+4j accumulator complete, 7+4j drained (warp 2 of each CTA)
#define NLA_STAMP3(slot) do { if (p.dbg && (slot) < 64) p.dbg[blockIdx.x * 64 + (slot)] = global_timer_ns(); } while (0)
  if (threadIdx.x == 0) NLA_STAMP3(0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; s++) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(smem_u32(&dfull_bar[b]), 1);
      mbar_init(smem_u32(&dfree_bar[b]), 2 * Shp::DRAIN_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc2(smem_u32(&tmem_slot), Shp::TMEM_COLS);
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&mapA); tma_prefetch_desc(&mapB); }
  tc_fence_before();
  __syncthreads();      // CTA-level ordering of the TMEM address written by tcgen05.alloc (the cluster barrier below covers it too, but
                        // compute-sanitizer racecheck only models CTA barriers for shared-memory hazards)
  cluster_sync_all();   // barriers of both CTAs initialised, TMEM allocated in both SMs
  tc_fence_after();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (tc_skip_launch(p)) ntiles = 0;   // conditioning guard: a rejected block-inverse leaf walks an empty tile list (uniform for the grid)
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) NLA_STAMP3(1);

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own rows of A, own half of the B tile; runs ahead across tile boundaries =====
    if (lane == 0) {
      uint32_t kc = 0;
      for (int r = 0; r * ncl < ntiles; r++) {
        const int i = tc3_index(r, cid, ncl);
        if (i >= ntiles) continue;
        const Tc3Tile t = tc3_tile(p, i, pairs_m);
        const int tm = 2 * t.pm + (int)rank;
        const int am = p.a_mn0 + tm * TC_BM, bn = p.b_mn0 + t.tn * BN + (int)rank * 128;
        const int nk = (t.klen + BK - 1) / BK;
        for (int kt = 0; kt < nk; kt++, kc++) {
          const int s = (int)(kc % S);
          const uint32_t it = kc / S;
          if (it > 0) mbar_wait_wd_cluster(smem_u32(&empty_bar[s]), (it - 1) & 1);
          const uint32_t sa = smem_base + s * STAGE, sb = sa + A_BYTES;
          const int ak = p.a_k0 + t.kbeg + kt * BK, bk = p.b_k0 + t.kbeg + kt * BK;
          const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);
          if (leader) mbar_expect_tx(smem_u32(&full_bar[s]), 2 * STAGE);
          if (AMAJ == MAJ_K) {
            tma_load_2d_pair(sa, &mapA, fb, ak, am);
          } else {
#pragma unroll
            for (int a = 0; a < TC_BM / ATOM; a++) tma_load_2d_pair(sa + a * (BK * 128), &mapA, fb, am + a * ATOM, ak);
          }
          if (BMAJ == MAJ_K) {
            tma_load_2d_pair(sb, &mapB, fb, bk, bn);
          } else {
#pragma unroll
            for (int a = 0; a < 128 / ATOM; a++) tma_load_2d_pair(sb + a * (BK * 128), &mapB, fb, bn + a * ATOM, bk);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===== MMA issuer of the pair =====
      constexpr uint32_t idesc = (1u << 4) | ((AMAJ == MAJ_MN ? 1u : 0u) << 15) | ((BMAJ == MAJ_MN ? 1u : 0u) << 16) |
                                 ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // F16 x F16 -> FP32, M = 256, N = 256
      constexpr uint32_t A_LBO = (AMAJ == MAJ_K) ? 16u : (uint32_t)(BK * 128), B_LBO = (BMAJ == MAJ_K) ? 16u : (uint32_t)(BK * 128);
      constexpr uint32_t A_KSTEP = (AMAJ == MAJ_K) ? 32u : (uint32_t)(UK * 128), B_KSTEP = (BMAJ == MAJ_K) ? 32u : (uint32_t)(UK * 128);
      constexpr uint32_t A_SBO = 1024u, B_SBO = 1024u;   // K-major: 8 rows of 128 B; MN-major: 8 k-rows of 128 B
      uint32_t kc = 0;
      int j = -1;
      for (int r = 0; r * ncl < ntiles; r++) {
        const int i = tc3_index(r, cid, ncl);
        if (i >= ntiles) continue;
        j++;
        const Tc3Tile t = tc3_tile(p, i, pairs_m);
        const int nk = (t.klen + BK - 1) / BK;
        const int buf = j & 1, use = j >> 1;
        if (use > 0) {   // the drain warps of both CTAs must have emptied this accumulator (tile j - 2)
          mbar_wait_wd_cluster(smem_u32(&dfree_bar[buf]), (use - 1) & 1);
          tc_fence_after();
        }
        const uint32_t dt = tmem + (uint32_t)(buf * BN);
        uint32_t acc = 0;
        for (int kt = 0; kt < nk; kt++, kc++) {
          const int s = (int)(kc % S);
          const uint32_t it = kc / S;
          mbar_wait_wd_cluster(smem_u32(&full_bar[s]), it & 1);
          tc_fence_after();
          if (kt == 0) NLA_STAMP3(4 + 4 * j);
          const uint32_t sa = smem_base + s * STAGE, sb = sa + A_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / UK; kk++) {
            const uint64_t da = umma_desc(sa + kk * A_KSTEP, A_LBO, A_SBO, UMMA_SW128);
            const uint64_t db = umma_desc(sb + kk * B_KSTEP, B_LBO, B_SBO, UMMA_SW128);
            tc_mma2_f16(dt, da, db, idesc, acc);
            acc = 1;
          }
          tc_commit2(smem_u32(&empty_bar[s]));
        }
        tc_commit2(smem_u32(&dfull_bar[buf]));
        NLA_STAMP3(5 + 4 * j);
      }
    }
    __syncwarp();
  } else {
    // ===== drain (both CTAs): 8 warps, warp quarter = warp & 3 (TMEM lane restriction), column half = (warp - 2) / 4 =====
    const int dw = warp - 2;
    const int quarter = warp & 3, half = dw >> 2;
    T* cbase = reinterpret_cast<T*>(p.C);
    float* stg = reinterpret_cast<float*>(smem_gen + S * STAGE) + dw * 1024;
    constexpr int VEC = 16 / ES, LPC = 32 / VEC, CPP = 32 / LPC, PASSES = 32 / CPP;   // 8 rows per lane, 4 lanes per column, 8 columns per pass
    const int rseg = (lane % LPC) * VEC;
    auto prefetch_tile = [&](const Tc3Tile& t) {   // pull the old C of a tile into L2 (each lane one 32-byte sector of its warp's 32 rows)
      const int row = (2 * t.pm + (int)rank) * TC_BM + quarter * 32 + lane;
      if (row < p.M && (lane % 16) == 0) {
        const int c_lo = t.tn * BN + half * 128, c_hi = min(p.N, c_lo + 128);
        for (int c = c_lo; c < c_hi; c++) asm volatile("prefetch.global.L2 [%0];" ::"l"(cbase + row + (long long)c * p.ldc));
      }
    };
    if (!p.overwrite && cid < ntiles) prefetch_tile(tc3_tile(p, cid, pairs_m));
    int j = -1;
#pragma unroll 1
    for (int r = 0; r * ncl < ntiles; r++) {
      const int i = tc3_index(r, cid, ncl);
      if (i >= ntiles) continue;
      j++;
      const Tc3Tile t = tc3_tile(p, i, pairs_m);
      const int tm = 2 * t.pm + (int)rank;
      const int buf = j & 1, use = j >> 1;
      if (!p.overwrite && tc3_index(r + 1, cid, ncl) < ntiles) prefetch_tile(tc3_tile(p, tc3_index(r + 1, cid, ncl), pairs_m));
      const int grow = tm * TC_BM + quarter * 32 + rseg;          // first of this lane's 8 rows (C coordinates)
      const int col0 = t.tn * BN + half * 128;                    // first column of this warp's half
      const int ncols = min(128, p.N - col0);                     // valid columns (<= 0: nothing to store)
      const bool need_old = !p.overwrite;
      // interior tiles (all 128 rows of the CTA and all 128 columns of the warp valid) take a compact vector path; edge tiles a
      // generic element-wise one kept out of line -- a fully unrolled ragged epilogue is 28 KB of code per 32-column slice, which
      // falls out of the instruction cache (measured: 10 us per tile)
      const bool fast = (tm * TC_BM + TC_BM <= p.M) && ncols == 128;
      // the old C of the whole tile (4 slices of 32 columns) is requested up front: the loads are in flight while the tile's last
      // MMAs run, instead of one L2 round trip per slice under a saturated L2
      uint4 old[4][PASSES];
#pragma unroll
      for (int q = 0; q < 4; q++)
#pragma unroll
        for (int ps = 0; ps < PASSES; ps++) old[q][ps] = make_uint4(0u, 0u, 0u, 0u);   // overwrite mode: the "old C" is zero
      if (need_old && fast) {
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
          for (int ps = 0; ps < PASSES; ps++)
            old[q][ps] = *reinterpret_cast<const uint4*>(cbase + grow + (long long)(col0 + 32 * q + ps * CPP + lane / LPC) * p.ldc);
      }
      mbar_wait_wd_cluster(smem_u32(&dfull_bar[buf]), use & 1);
      tc_fence_after();
      if (threadIdx.x == 64) NLA_STAMP3(6 + 4 * j);
      const uint32_t dt = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN + half * 128);
      if (fast) {
        // dup destination of this lane's rows (block-inverse leaves): in range for all 8 rows or for none (ranges are multiples of 128)
        const int dr = grow - p.dup_r0;
        const bool dup_rows = p.dup != nullptr && dr >= 0 && dr + VEC <= p.dup_rn;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          uint32_t r[32];
          tmem_ld32(dt + (uint32_t)(32 * q), r);
          tmem_ld_wait();
          __syncwarp();   // the previous slice has been read out of the staging tile
#pragma unroll
          for (int c = 0; c < 32; c++) stg[c * 32 + lane] = __uint_as_float(r[c]);
          __syncwarp();
#pragma unroll
          for (int ps = 0; ps < PASSES; ps++) {
            const int cl = ps * CPP + lane / LPC, col = col0 + 32 * q + cl;
            const float4 a0 = *reinterpret_cast<const float4*>(&stg[cl * 32 + rseg]), a1 = *reinterpret_cast<const float4*>(&stg[cl * 32 + rseg + 4]);
            const float accv[VEC] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const T* ov = reinterpret_cast<const T*>(&old[q][ps]);
            T outv[VEC];
#pragma unroll
            for (int e = 0; e < VEC; e++) {
              // rounding points as in gemm_tc.cuh: beta (alpha folded into the first touch) rounds to T, one rounding of the update,
              // post rounds again.  Branch-free on purpose: for beta = 1 / post = 1 the extra roundings are exact no-ops (the value
              // already is a Float16 number / rounding is idempotent), and a branchy version makes the compiler clone this block
              // eight times per pass (120 KB of code, instruction-cache bound).
              float v = __half2float(__float2half_rn(p.beta * __half2float(ov[e])));
              v = fmaf(p.sgn, accv[e], v);
              outv[e] = __float2half_rn(p.post * __half2float(__float2half_rn(v)));
            }
            *reinterpret_cast<uint4*>(cbase + grow + (long long)col * p.ldc) = *reinterpret_cast<const uint4*>(outv);
            const int dc = col - p.dup_c0;
            if (dup_rows && dc >= 0 && dc < p.dup_cn)
              *reinterpret_cast<uint4*>(reinterpret_cast<T*>(p.dup) + dr + (long long)dc * p.dup_ld) = *reinterpret_cast<const uint4*>(outv);
          }
        }
      } else {
#pragma unroll 1
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(dt + (uint32_t)c0, r);
          tmem_ld_wait();
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 32; c++) stg[c * 32 + lane] = __uint_as_float(r[c]);
          __syncwarp();
          tc3_store_edge(p, stg, lane, tm * TC_BM + quarter * 32, col0 + c0, min(32, ncols - c0));
        }
      }
      // hand the accumulator back to the MMA issuer of the pair
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&dfree_bar[buf]), 0));
      if (threadIdx.x == 64) NLA_STAMP3(7 + 4 * j);
    }
  }

  __syncwarp();   // the cluster barrier below is warp-aligned: reconverge the single-lane roles first
  // ===== teardown: nobody may leave (or free TMEM) while the peer still uses this CTA's shared memory / tensor memory =====
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(tmem, Shp::TMEM_COLS);
#undef NLA_STAMP3
}

}  // namespace nla

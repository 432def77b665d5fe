// gemm_tc4.cuh -- persistent CTA-pair kernel with 256 x 512 tiles for the LONG Float16 updates (K >= 4096).
//
// Same operation and epilogue arithmetic as gemm_tc3.cuh (C <- post*(beta*C + sgn*opA(A)*opB(B)), src/matmul.jl:5-81 in the
// reference).  The 256 x 256 pair tile moves 64 KB of operands from L2 per K step of 64 for 8.4 MFLOP = 128 flop per byte, and
// the in-kernel stamps of gemm_tc3 put the chip's L2 throughput (10-11 TB/s of operand traffic with all 74 pairs streaming) at
// exactly the 1.4 PFLOP/s that the long updates reach.  Here a pair computes TWO 256 x 256 accumulators that share
// the A tile: per K step each CTA stages its 128 rows of A once and 2 x 128 columns of B -- 96 KB per pair for 16.8 MFLOP,
// 171 flop per byte of L2 traffic.  The price: both accumulators fill tensor memory (2 x 256 columns), so there is no second
// buffer and the drain of a tile is exposed (~10 us) before the next tile's first MMA -- negligible against a main loop of
// 64+ K steps, which is why only the long updates come here; 4-stage ring (4 x 48 KB) + 32 KB of drain staging per CTA.
#pragma once
#include "gemm_tc3.cuh"

namespace nla {

struct Tc4Shape {
  static constexpr int BK = 64, UK = 16;
  static constexpr int A_BYTES = TC_BM * BK * 2;          // own 128 rows of A
  static constexpr int B_SUB = 128 * BK * 2;              // own 128 columns of ONE of the two accumulators' B tiles
  static constexpr int STAGE = A_BYTES + 2 * B_SUB;       // 48 KB
  static constexpr int STAGES = 4;
  static constexpr int DRAIN_WARPS = 8;
  static constexpr int STG_BYTES = DRAIN_WARPS * 4096;
  static constexpr int SMEM = STAGES * STAGE + STG_BYTES + 1024;
  static constexpr int THREADS = 64 + 32 * DRAIN_WARPS;
  static constexpr int BN = 512;
  static constexpr int TMEM_COLS = 512;
};

// pair-tile i -> (pm, tn): groups of 4 pairs of M tiles share a B tile, as in the other kernels
__device__ __forceinline__ void tc4_tile(const GemmTcParams& p, int i, int pairs_m, int& pm, int& tn) {
  constexpr int GROUP_P = TC_GROUP_M / 2;
  const int per_group = GROUP_P * p.tiles_n;
  const int grp = i / per_group;
  const int first_p = grp * GROUP_P;
  const int gsz = min(GROUP_P, pairs_m - first_p);
  const int rem = i - grp * per_group;
  pm = first_p + rem % gsz;
  tn = rem / gsz;
}

template <int AMAJ, int BMAJ>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Tc4Shape::THREADS, 1)
gemm_tc4_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const GemmTcParams p) {
  using T = __half;
  using Shp = Tc4Shape;
  constexpr int S = Shp::STAGES, BK = Shp::BK, UK = Shp::UK, BN = Shp::BN;
  constexpr int ES = 2, ATOM = 128 / ES;
  constexpr int A_BYTES = Shp::A_BYTES, B_SUB = Shp::B_SUB, STAGE = Shp::STAGE;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[S];      // leader: the bytes of BOTH CTAs have landed
  __shared__ __align__(8) uint64_t empty_bar[S];     // local: the pair's MMAs that read this stage have completed (multicast commit)
  __shared__ __align__(8) uint64_t dfull_bar;        // local: the tile has been accumulated (multicast commit)
  __shared__ __align__(8) uint64_t dfree_bar;        // leader: both CTAs have drained tensor memory (one arrival per drain warp)
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pairs_m = (p.tiles_m + 1) >> 1;
  const int ntiles = pairs_m * p.tiles_n;            // tiles_n counts 512-wide tiles
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; s++) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&dfull_bar), 1);
    mbar_init(smem_u32(&dfree_bar), 2 * Shp::DRAIN_WARPS);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc2(smem_u32(&tmem_slot), Shp::TMEM_COLS);
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&mapA); tma_prefetch_desc(&mapB); }
  tc_fence_before();
  __syncthreads();      // CTA-level ordering of the TMEM address written by tcgen05.alloc (the cluster barrier below covers it too, but
                        // compute-sanitizer racecheck only models CTA barriers for shared-memory hazards)
  cluster_sync_all();
  tc_fence_after();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int nk = (p.K + BK - 1) / BK;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own rows of A, own 128 columns of each accumulator's B tile =====
    if (lane == 0) {
      uint32_t kc = 0;
      for (int i = cid; i < ntiles; i += ncl) {
        int pm, tn;
        tc4_tile(p, i, pairs_m, pm, tn);
        const int tm = 2 * pm + (int)rank;
        const int am = p.a_mn0 + tm * TC_BM;
        const int bn0 = p.b_mn0 + tn * BN + (int)rank * 128, bn1 = bn0 + 256;
        for (int kt = 0; kt < nk; kt++, kc++) {
          const int s = (int)(kc % S);
          const uint32_t it = kc / S;
          if (it > 0) mbar_wait_wd_cluster(smem_u32(&empty_bar[s]), (it - 1) & 1);
          const uint32_t sa = smem_base + s * STAGE, sb0 = sa + A_BYTES, sb1 = sb0 + B_SUB;
          const int ak = p.a_k0 + kt * BK, bk = p.b_k0 + kt * BK;
          const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);
          if (leader) mbar_expect_tx(smem_u32(&full_bar[s]), 2 * STAGE);
          if (AMAJ == MAJ_K) {
            tma_load_2d_pair(sa, &mapA, fb, ak, am);
          } else {
#pragma unroll
            for (int a = 0; a < TC_BM / ATOM; a++) tma_load_2d_pair(sa + a * (BK * 128), &mapA, fb, am + a * ATOM, ak);
          }
          if (BMAJ == MAJ_K) {
            tma_load_2d_pair(sb0, &mapB, fb, bk, bn0);
            tma_load_2d_pair(sb1, &mapB, fb, bk, bn1);
          } else {
#pragma unroll
            for (int a = 0; a < 128 / ATOM; a++) tma_load_2d_pair(sb0 + a * (BK * 128), &mapB, fb, bn0 + a * ATOM, bk);
#pragma unroll
            for (int a = 0; a < 128 / ATOM; a++) tma_load_2d_pair(sb1 + a * (BK * 128), &mapB, fb, bn1 + a * ATOM, bk);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===== MMA issuer of the pair: two M256 x N256 accumulators (TMEM columns 0 and 256) share every A tile =====
      constexpr uint32_t idesc = (1u << 4) | ((AMAJ == MAJ_MN ? 1u : 0u) << 15) | ((BMAJ == MAJ_MN ? 1u : 0u) << 16) |
                                 ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      constexpr uint32_t A_LBO = (AMAJ == MAJ_K) ? 16u : (uint32_t)(BK * 128), B_LBO = (BMAJ == MAJ_K) ? 16u : (uint32_t)(BK * 128);
      constexpr uint32_t A_KSTEP = (AMAJ == MAJ_K) ? 32u : (uint32_t)(UK * 128), B_KSTEP = (BMAJ == MAJ_K) ? 32u : (uint32_t)(UK * 128);
      uint32_t kc = 0;
      int j = 0;
      for (int i = cid; i < ntiles; i += ncl, j++) {
        if (j > 0) {   // single accumulator set: the previous tile must have been drained by both CTAs
          mbar_wait_wd_cluster(smem_u32(&dfree_bar), (j - 1) & 1);
          tc_fence_after();
        }
        uint32_t acc = 0;
        for (int kt = 0; kt < nk; kt++, kc++) {
          const int s = (int)(kc % S);
          const uint32_t it = kc / S;
          mbar_wait_wd_cluster(smem_u32(&full_bar[s]), it & 1);
          tc_fence_after();
          const uint32_t sa = smem_base + s * STAGE, sb0 = sa + A_BYTES, sb1 = sb0 + B_SUB;
#pragma unroll
          for (int kk = 0; kk < BK / UK; kk++) {
            const uint64_t da = umma_desc(sa + kk * A_KSTEP, A_LBO, 1024u, UMMA_SW128);
            const uint64_t db0 = umma_desc(sb0 + kk * B_KSTEP, B_LBO, 1024u, UMMA_SW128);
            const uint64_t db1 = umma_desc(sb1 + kk * B_KSTEP, B_LBO, 1024u, UMMA_SW128);
            tc_mma2_f16(tmem, da, db0, idesc, acc);
            tc_mma2_f16(tmem + 256u, da, db1, idesc, acc);
            acc = 1;
          }
          tc_commit2(smem_u32(&empty_bar[s]));
        }
        tc_commit2(smem_u32(&dfull_bar));
      }
    }
    __syncwarp();
  } else {
    // ===== drain (both CTAs): 8 warps; warp quarter = warp & 3 (TMEM lanes), accumulator = (warp - 2) / 4, two groups of 128 columns each =====
    const int dw = warp - 2;
    const int quarter = warp & 3, half = dw >> 2;
    T* cbase = reinterpret_cast<T*>(p.C);
    float* stg = reinterpret_cast<float*>(smem_gen + S * STAGE) + dw * 1024;
    constexpr int VEC = 16 / ES, LPC = 32 / VEC, CPP = 32 / LPC, PASSES = 32 / CPP;
    const int rseg = (lane % LPC) * VEC;
    const bool need_old = !p.overwrite;
    int j = 0;
#pragma unroll 1
    for (int i = cid; i < ntiles; i += ncl, j++) {
      int pm, tn;
      tc4_tile(p, i, pairs_m, pm, tn);
      const int tm = 2 * pm + (int)rank;
      const int grow = tm * TC_BM + quarter * 32 + rseg;
      bool waited = false;
#pragma unroll 1
      for (int g = 0; g < 2; g++) {
        const int col0 = tn * BN + half * 256 + g * 128;      // first column of this group (C coordinates)
        const int ncols = min(128, p.N - col0);
        const bool fast = (tm * TC_BM + TC_BM <= p.M) && ncols == 128;
        uint4 old[4][PASSES];
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
          for (int ps = 0; ps < PASSES; ps++) old[q][ps] = make_uint4(0u, 0u, 0u, 0u);
        if (need_old && fast) {   // requested before the accumulator is waited for (first group) / while nothing else is pending (second)
#pragma unroll
          for (int q = 0; q < 4; q++)
#pragma unroll
            for (int ps = 0; ps < PASSES; ps++)
              old[q][ps] = *reinterpret_cast<const uint4*>(cbase + grow + (long long)(col0 + 32 * q + ps * CPP + lane / LPC) * p.ldc);
        }
        if (!waited) {
          mbar_wait_wd_cluster(smem_u32(&dfull_bar), j & 1);
          tc_fence_after();
          waited = true;
        }
        const uint32_t dt = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 256 + g * 128);
        const int dr = grow - p.dup_r0;   // second destination (the next block-inverse leaf's copy of V): all 8 rows of a lane or none
        const bool dup_rows = p.dup != nullptr && dr >= 0 && dr + VEC <= p.dup_rn;
        if (fast) {
#pragma unroll
          for (int q = 0; q < 4; q++) {
            uint32_t r[32];
            tmem_ld32(dt + (uint32_t)(32 * q), r);
            tmem_ld_wait();
            __syncwarp();
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
              for (int e = 0; e < VEC; e++) {   // branch-free, same rounding points as gemm_tc3.cuh
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
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&dfree_bar), 0));
    }
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(tmem, Shp::TMEM_COLS);
}

}  // namespace nla

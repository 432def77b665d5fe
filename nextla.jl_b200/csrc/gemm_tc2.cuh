// gemm_tc2.cuh -- the large Float16 / Float32 updates on CTA PAIRS:  tcgen05.mma.cta_group::2, M = 256 per instruction.
//
// Same operation, operands, TMA maps and drain as gemm_tc.cuh (C <- post*(beta*C + sgn*opA(A)*opB(B)), src/matmul.jl:5-81 in the
// reference), different machine mapping.  ncu on the single-CTA kernel (profiles/r01_ncu_prof_tc_f16_top_summary.csv) shows
// the tensor pipe 71 % active with DRAM at 23 %: every CTA streams its own A (128 x K) AND a full B (256 x K) tile through L2,
// 96 B/clk/SM at full tensor rate.  A pair of CTAs on the two SMs of one TPC computes a 256 x 256 tile with ONE instruction
// stream: each CTA stages its own 128 rows of A and only HALF of the B tile (128 of the 256 columns); the tensor cores of both
// SMs read both halves.  Per SM that is 2/3 of the L2 -> shared-memory traffic and 2/3 of the shared-memory operand reads
// (for Float32 the hi/lo splitter work per SM drops by the same third).
//
// Protocol (rank 0 = leader):
//   * every CTA: TMA producer warp (own A rows, own half of B) with local full/empty mbarriers -- exactly as in gemm_tc.cuh;
//   * Float16: both CTAs' TMA copies complete on the LEADER's full barrier (cp.async.bulk.tensor.cta_group::2);
//     Float32: TMA completes locally (the splitter warps of each CTA need it), then one lane per splitter warp of BOTH CTAs
//     arrives on the leader's conv barrier (a first version with 128 remote arrivals per CTA and stage ran 40 % slower);
//   * the leader's MMA thread issues tcgen05.mma.cta_group::2 and commits with multicast: the empty barrier of the stage and
//     the accumulator-full barrier fire in both CTAs;
//   * both CTAs drain their own 128 rows x 256 columns from their own TMEM; for Float32 (K chunks, two TMEM tiles) both sets
//     of drain warps arrive on the leader's dfree barrier (count 256);
//   * cluster barriers after the prologue (remote arrivals must find initialised barriers) and before TMEM release.
#pragma once
#include "gemm_tc.cuh"

namespace nla {

template <typename T> struct Tc2Shape {
  static constexpr int A_BYTES = TC_BM * TcCfg<T>::BK * (int)sizeof(T);   // own 128 rows of A
  static constexpr int B_BYTES = 128 * TcCfg<T>::BK * (int)sizeof(T);     // own 128 of the pair's 256 columns of B
  static constexpr int HALF_STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGE = HALF_STAGE * (TcCfg<T>::PASSES == 3 ? 2 : 1);
  static constexpr int STAGES = (sizeof(T) == 4) ? 6 : 3;                 // Float16: 3 x 32 KB (2 CTAs/SM); Float32: 6 x 32 KB
  static constexpr int SMEM = STAGES * STAGE + 1024 + (TcCfg<T>::PASSES == 3 ? 16384 : 0);
  static constexpr int BN = 256;
};

template <typename T, int AMAJ, int BMAJ>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TcCfg<T>::THREADS, TcCfg<T>::MIN_CTAS)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const GemmTcParams p) {
  using Cfg = TcCfg<T>;
  using Shp = Tc2Shape<T>;
  constexpr int S = Shp::STAGES, BK = Cfg::BK, UK = Cfg::UK, BN = Shp::BN;
  constexpr int ES = (int)sizeof(T);
  constexpr int ATOM = 128 / ES;
  constexpr int A_BYTES = Shp::A_BYTES, HALF_STAGE = Shp::HALF_STAGE, STAGE = Shp::STAGE;
  constexpr bool F32 = Cfg::PASSES == 3;
  constexpr int NBUF = Cfg::NBUF;
  constexpr int TMEM_COLS = NBUF * BN;
  constexpr int DRAIN_WARP0 = Cfg::THREADS / 32 - 4;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[S];       // Float32: my own TMA bytes have landed; Float16: (leader) the bytes of BOTH CTAs have landed
  __shared__ __align__(8) uint64_t conv_bar[S];       // leader: lo tiles of both CTAs written (one arrival per splitter warp, Float32)
  __shared__ __align__(8) uint64_t empty_bar[S];      // local: the pair's MMAs that read this stage have completed (multicast commit)
  __shared__ __align__(8) uint64_t dfull_bar[NBUF];   // local: a K chunk has been accumulated (multicast commit)
  __shared__ __align__(8) uint64_t dfree_bar[NBUF];   // leader: both CTAs have drained this TMEM tile (one arrival per drain warp)
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  // grouped rasterisation over PAIRS of M tiles
  const int pairs_m = (p.tiles_m + 1) >> 1;
  const int pair_id = blockIdx.x >> 1;
  constexpr int GROUP_P = TC_GROUP_M / 2;
  const int per_group = GROUP_P * p.tiles_n;
  const int grp = pair_id / per_group;
  const int first_p = grp * GROUP_P;
  const int gsz = min(GROUP_P, pairs_m - first_p);
  const int rem = pair_id - grp * per_group;
  const int pm = first_p + rem % gsz;
  const int tn = rem / gsz;
  const int tm = 2 * pm + (int)rank;   // may be one past the last M tile: that CTA stages zeros and stores nothing

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; s++) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&conv_bar[s]), 8);          // one arrival per splitter warp of either CTA
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < NBUF; b++) {
      mbar_init(smem_u32(&dfull_bar[b]), 1);
      mbar_init(smem_u32(&dfree_bar[b]), 8);         // one arrival per drain warp of either CTA
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc2(smem_u32(&tmem_slot), TMEM_COLS);
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&mapA); tma_prefetch_desc(&mapB); }
  tc_fence_before();
  __syncthreads();      // CTA-level ordering of the TMEM address written by tcgen05.alloc (the cluster barrier below covers it too, but
                        // compute-sanitizer racecheck only models CTA barriers for shared-memory hazards)
  cluster_sync_all();   // barriers of both CTAs initialised, TMEM allocated in both SMs
  tc_fence_after();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int nk = (p.K + BK - 1) / BK;
  const int chunk_kb = (NBUF > 1 && p.chunk_k > 0) ? max(1, p.chunk_k / BK) : nk;
  const int nchunks = (nk + chunk_kb - 1) / chunk_kb;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own rows of A, own half of the B tile =====
    if (lane == 0) {
      const int am = p.a_mn0 + tm * TC_BM, bn = p.b_mn0 + tn * BN + (int)rank * 128;
      for (int kt = 0; kt < nk; kt++) {
        const int s = kt % S, it = kt / S;
        if (it > 0) mbar_wait_wd_cluster(smem_u32(&empty_bar[s]), (it - 1) & 1);
        const uint32_t sa = smem_base + s * STAGE, sb = sa + A_BYTES;
        const int ak = p.a_k0 + kt * BK, bk = p.b_k0 + kt * BK;
        if (F32) {
          // the splitter warps of this CTA wait for the bytes: local barrier, plain TMA
          const uint32_t fb = smem_u32(&full_bar[s]);
          mbar_expect_tx(fb, HALF_STAGE);
          if (AMAJ == MAJ_K) {
            tma_load_2d(sa, &mapA, fb, ak, am);
          } else {
#pragma unroll
            for (int a = 0; a < TC_BM / ATOM; a++) tma_load_2d(sa + a * (BK * 128), &mapA, fb, am + a * ATOM, ak);
          }
          if (BMAJ == MAJ_K) {
            tma_load_2d(sb, &mapB, fb, bk, bn);   // box {BK, 128}
          } else {
#pragma unroll
            for (int a = 0; a < 128 / ATOM; a++) tma_load_2d(sb + a * (BK * 128), &mapB, fb, bn + a * ATOM, bk);
          }
        } else {
          // Float16: both CTAs' copies complete on the LEADER's barrier (cta_group::2 TMA); the leader arms it for both halves
          const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);
          if (leader) mbar_expect_tx(smem_u32(&full_bar[s]), 2 * HALF_STAGE);
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
      constexpr uint32_t idesc = (1u << 4) | (Cfg::FMT << 7) | (Cfg::FMT << 10) | ((AMAJ == MAJ_MN ? 1u : 0u) << 15) |
                                 ((BMAJ == MAJ_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      constexpr uint32_t A_LBO = (AMAJ == MAJ_K) ? 16u : (uint32_t)(BK * 128), B_LBO = (BMAJ == MAJ_K) ? 16u : (uint32_t)(BK * 128);
      constexpr uint32_t A_KSTEP = (AMAJ == MAJ_K) ? 32u : (uint32_t)(UK * 128), B_KSTEP = (BMAJ == MAJ_K) ? 32u : (uint32_t)(UK * 128);
      constexpr bool A32 = (AMAJ == MAJ_MN) && ES == 4, B32 = (BMAJ == MAJ_MN) && ES == 4;
      constexpr uint32_t KROW = (uint32_t)(BK * ES);
      constexpr uint32_t K_LAY = (KROW == 128) ? UMMA_SW128 : UMMA_SW64;
      constexpr uint32_t A_SBO = (AMAJ == MAJ_K) ? 8u * KROW : (A32 ? 512u : 1024u), B_SBO = (BMAJ == MAJ_K) ? 8u * KROW : (B32 ? 512u : 1024u);
      constexpr uint32_t A_LAY = (AMAJ == MAJ_K) ? K_LAY : (A32 ? UMMA_SW128_BASE32B : UMMA_SW128);
      constexpr uint32_t B_LAY = (BMAJ == MAJ_K) ? K_LAY : (B32 ? UMMA_SW128_BASE32B : UMMA_SW128);
      int kt = 0;
      for (int c = 0; c < nchunks; c++) {
        const int buf = c % NBUF, use = c / NBUF;
        if (use > 0) {
          mbar_wait_wd_cluster(smem_u32(&dfree_bar[buf]), (use - 1) & 1);
          tc_fence_after();
        }
        const uint32_t dt = tmem + (uint32_t)(buf * BN);
        uint32_t acc = 0;
        const int kend = min(nk, kt + chunk_kb);
        for (; kt < kend; kt++) {
          const int s = kt % S, it = kt / S;
          mbar_wait_wd_cluster(smem_u32(F32 ? &conv_bar[s] : &full_bar[s]), it & 1);
          tc_fence_after();
          const uint32_t sa = smem_base + s * STAGE, sb = sa + A_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / UK; kk++) {
            const uint64_t da = umma_desc(sa + kk * A_KSTEP, A_LBO, A_SBO, A_LAY);
            const uint64_t db = umma_desc(sb + kk * B_KSTEP, B_LBO, B_SBO, B_LAY);
            if (!F32) {
              tc_mma2_f16(dt, da, db, idesc, acc);
            } else {
              const uint64_t dal = umma_desc(sa + HALF_STAGE + kk * A_KSTEP, A_LBO, A_SBO, A_LAY);
              const uint64_t dbl = umma_desc(sb + HALF_STAGE + kk * B_KSTEP, B_LBO, B_SBO, B_LAY);
              tc_mma2_tf32(dt, dal, db, idesc, acc);   // lo * hi
              tc_mma2_tf32(dt, da, dbl, idesc, 1u);    // hi * lo
              tc_mma2_tf32(dt, da, db, idesc, 1u);     // hi * hi
            }
            acc = 1;
          }
          tc_commit2(smem_u32(&empty_bar[s]));
        }
        tc_commit2(smem_u32(&dfull_bar[buf]));
      }
    }
    __syncwarp();
  } else if (warp < DRAIN_WARP0) {
    // ===== Float32: write the lo tile of my operands, then arrive on the LEADER's conv barrier =====
    const int et = threadIdx.x - 64;
    const uint32_t hi_round = p.raw_hi == 0 ? 0x1000u : 0u, lo_round = p.raw_hi == 2 ? 0u : 0x1000u;
    for (int kt = 0; kt < nk; kt++) {
      const int s = kt % S, it = kt / S;
      mbar_wait_wd(smem_u32(&full_bar[s]), it & 1);
      uint4* hi = reinterpret_cast<uint4*>(smem_gen + s * STAGE);
      uint4* lo = reinterpret_cast<uint4*>(smem_gen + s * STAGE + HALF_STAGE);
#pragma unroll 4
      for (int i = et; i < HALF_STAGE / 16; i += 128) {
        uint4 v = hi[i], h, l;
        tf32_split(v.x, hi_round, lo_round, h.x, l.x); tf32_split(v.y, hi_round, lo_round, h.y, l.y);
        tf32_split(v.z, hi_round, lo_round, h.z, l.z); tf32_split(v.w, hi_round, lo_round, h.w, l.w);
        if (hi_round) hi[i] = h;
        lo[i] = l;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&conv_bar[s]), 0));   // one remote arrival per warp
    }
  } else {
    // ===== drain (both CTAs): own 128 rows x 256 columns =====
    const int quarter = warp & 3;
    const int row = tm * TC_BM + quarter * 32 + lane;
    const bool row_ok = row < p.M;
    T* cbase = reinterpret_cast<T*>(p.C);
    T* crow = cbase + row;
    const int ncols = min(BN, p.N - tn * BN);
    float* stg = reinterpret_cast<float*>(smem_gen + (F32 ? S * STAGE : 0)) + (warp - DRAIN_WARP0) * 1024;
    if (!p.overwrite) {
      if (row_ok && (lane % (32 / ES)) == 0)
        for (int j = 0; j < ncols; j++) asm volatile("prefetch.global.L2 [%0];" ::"l"(crow + (long long)(tn * BN + j) * p.ldc));
    }
#pragma unroll 1
    for (int c = 0; c < nchunks; c++) {
      const int buf = c % NBUF, use = c / NBUF;
      mbar_wait_wd_cluster(smem_u32(&dfull_bar[buf]), use & 1);
      tc_fence_after();
      const bool first = (c == 0), last = (c == nchunks - 1);
      const bool need_old = !(first && p.overwrite);
      const float beta = first ? p.beta : 1.0f, post = last ? p.post : 1.0f;
      const uint32_t dt = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (c0 >= ncols) break;
        uint32_t r[32];
        tmem_ld32(dt + (uint32_t)c0, r);
        tmem_ld_wait();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j++) stg[j * 32 + lane] = __uint_as_float(r[j]);
        __syncwarp();
        constexpr int VEC = 16 / ES, LPC = 32 / VEC, CPP = 32 / LPC, PASSES = 32 / CPP;
        const int rseg = (lane % LPC) * VEC;
        const int grow = tm * TC_BM + quarter * 32 + rseg;
        uint4 oldv[PASSES];
#pragma unroll
        for (int ps = 0; ps < PASSES; ps++) oldv[ps] = make_uint4(0u, 0u, 0u, 0u);
        if (need_old) {
#pragma unroll
          for (int ps = 0; ps < PASSES; ps++) {
            const int col = c0 + ps * CPP + lane / LPC;
            oldv[ps] = make_uint4(0u, 0u, 0u, 0u);
            if (col < ncols && grow + VEC <= p.M) {
              oldv[ps] = *reinterpret_cast<const uint4*>(cbase + grow + (long long)(tn * BN + col) * p.ldc);
            } else if (col < ncols && grow < p.M) {
              T tmp[VEC];
#pragma unroll
              for (int e = 0; e < VEC; e++) tmp[e] = (grow + e < p.M) ? cbase[grow + e + (long long)(tn * BN + col) * p.ldc] : tc_from_float<T>(0.f);
              oldv[ps] = *reinterpret_cast<const uint4*>(tmp);
            }
          }
        }
#pragma unroll
        for (int ps = 0; ps < PASSES; ps++) {
          const int cl = ps * CPP + lane / LPC, col = c0 + cl;
          if (col < ncols && grow < p.M) {
            T outv[VEC];
            const T* ov = reinterpret_cast<const T*>(&oldv[ps]);
            float accv[VEC];
#pragma unroll
            for (int e = 0; e < VEC; e += 4) *reinterpret_cast<float4*>(&accv[e]) = *reinterpret_cast<const float4*>(&stg[cl * 32 + rseg + e]);
#pragma unroll
            for (int e = 0; e < VEC; e++) {
              // branch-free: for beta = 1 / post = 1 the extra roundings are exact no-ops, and a branchy version is cloned by the
              // compiler for every combination (instruction-cache pressure in the drain, see gemm_tc3.cuh); oldv is zero when unused
              float v = tc_to_float<T>(tc_from_float<T>(beta * tc_to_float<T>(ov[e])));
              v += p.sgn * accv[e];
              outv[e] = tc_from_float<T>(post * tc_to_float<T>(tc_from_float<T>(v)));
            }
            T* dst = cbase + grow + (long long)(tn * BN + col) * p.ldc;
            if (grow + VEC <= p.M) {
              *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(outv);
            } else {
#pragma unroll
              for (int e = 0; e < VEC; e++) if (grow + e < p.M) dst[e] = outv[e];
            }
            if (last && p.dup) {   // second copy of the rows/columns the next block-inverse leaf reads (its GEMM is out of place)
              const int dr = grow - p.dup_r0, dc = tn * BN + col - p.dup_c0;
              if (dr >= 0 && dr < p.dup_rn && dc >= 0 && dc < p.dup_cn) {
                T* dd = reinterpret_cast<T*>(p.dup) + dr + (long long)dc * p.dup_ld;
                if (grow + VEC <= p.M && dr + VEC <= p.dup_rn) {
                  *reinterpret_cast<uint4*>(dd) = *reinterpret_cast<const uint4*>(outv);
                } else {
#pragma unroll
                  for (int e = 0; e < VEC; e++) if (grow + e < p.M && dr + e < p.dup_rn) dd[e] = outv[e];
                }
              }
            }
          }
        }
      }
      if (NBUF > 1 && !last) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&dfree_bar[buf]), 0));
      }
    }
  }

  __syncwarp();   // the cluster barrier below is warp-aligned: reconverge the single-lane roles first
  // ===== teardown: nobody may leave (or free TMEM) while the peer still uses this CTA's shared memory / tensor memory =====
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(tmem, TMEM_COLS);
}

}  // namespace nla

// gemm_tc.cuh -- Float16 / Float32 Schur/GEMM update on the 5th-generation tensor cores (tcgen05, accumulators in TMEM):
//     C <- post * (beta*C + sgn * opA(A) * opB(B))          (update)
//     C <- scale * opA(A) * opB(B)                          (overwrite: the diagonal-block leaves, see diag_prep.cuh)
//
// Replaces the reference's `matmul!` KernelAbstractions kernel (src/matmul.jl:5-66, launched from GEMM_ADD!/GEMM_SUB!
// :69-81; one output per work-item, accumulation in the element type) for the two low-precision element types.
//
// B200 design:
//  * CTA tile 128 x 256 (or 128 x 128), one `tcgen05.mma.cta_group::1` of shape M128 x N256 x K(32 bytes) per step, issued by ONE thread;
//    the FP32 accumulator tile (128 lanes x 256 columns) lives in tensor memory.
//  * Float16: kind::f16 (FP16 inputs, FP32 accumulate -- north star; the reference accumulates in FP16).
//  * Float32: kind::tf32 three times per step ("3xTF32"): every operand tile is split in shared memory into
//    hi = a & 0xffffe000 (exactly a TF32 number; the tensor core does this truncation itself, so the raw tile serves as hi)
//    and lo = a - hi (exact in FP32), and D += hi*hi + hi*lo + lo*hi.  The
//    dropped lo*lo term and the TF32 rounding of lo leave a relative error of ~2^-21 per product, inside the reference's
//    1e-5 Float32 tolerance (test/trsm.jl:8); a single TF32 pass (2^-11) would not be.
//  * Operands come straight from the caller's column-major matrices through 2-D TMA tensor maps with SWIZZLE_128B:
//      K-major operand  (K contiguous in memory):  one box {128 B of K, 128|256 rows}
//      MN-major operand (M/N contiguous)        :  boxes {128 B of M/N, BK rows of K}, one per 128-byte M/N atom
//    and the tcgen05 shared-memory descriptors describe exactly those layouts (see umma_desc in common.cuh).  MN-major
//    Float32 operands use the 32-byte-chunk variant of the swizzle on both sides (the only one tcgen05 takes for them).
//  * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, (Float32: warps 2-5 = hi/lo splitters,)
//    last 4 warps = drain/epilogue (TMEM -> registers -> C).  mbarrier rings between them; `tcgen05.commit` releases a
//    stage back to the producer and hands a finished accumulator chunk to the drain warps.
//  * Float32 accumulates K in chunks of 512 into two alternating TMEM tiles; the drain warps add each finished chunk into C
//    with round-to-nearest while the next chunk runs.  The tensor core's internal accumulation truncates (measured: error
//    grows linearly with the number of MMAs, ~2^-24 relative each, all in one direction for same-sign data), so one
//    8192-long accumulation would cost ~1e-4; chunking keeps the Float32 path at ~1e-6 in that worst case.
//  * Float16 stages are 48 KB (2 stages, 2 CTAs per SM so one CTA's epilogue overlaps the other's main loop);
//    Float32 stages are 96 KB (raw/hi + lo copies, 2 stages, 1 CTA per SM).  A 128 x 128 tile variant (3 stages) fills the
//    machine when the 256-wide grid would not.
//  * Epilogue: the old C tile is prefetched into L2 at kernel start; 32 columns (tcgen05.ld.32x32b.x32) per iteration go
//    through a per-warp FP32 staging tile in shared memory so that C is read and written with 16-byte accesses (element-wise
//    2-byte stores straight from the TMEM row layout cost 9 us per 128 x 128 tile, measured with in-kernel timestamps).
#pragma once
#include "common.cuh"
#include "gemm_f64.cuh"  // MAJ_MN / MAJ_K

namespace nla {

constexpr int TC_BM = 128;
constexpr int TC_GROUP_M = 8;

template <typename T> struct TcCfg;
template <> struct TcCfg<__half> {
  static constexpr int BK = 64, UK = 16, PASSES = 1, MIN_CTAS = 2;
  static constexpr int THREADS = 192;   // TMA warp, MMA warp, 4 epilogue warps
  static constexpr int NBUF = 1, CHUNK_K = 0;   // one accumulator tile, the whole K range in one chunk
  static constexpr uint32_t FMT = 0;  // F16
};
template <> struct TcCfg<float> {
  static constexpr int BK = 16, UK = 8, PASSES = 3, MIN_CTAS = 1;   // 64-byte K rows: half-size stages, twice the pipeline depth
  static constexpr int THREADS = 320;   // TMA warp, MMA warp, 4 hi/lo splitter warps, 4 drain warps
  static constexpr int NBUF = 2, CHUNK_K = 512;   // two accumulator tiles; every 512 of K is promoted into C with round-to-nearest
  static constexpr uint32_t FMT = 2;  // TF32
};

// Tile shapes: BN = 256 is the throughput shape (one MMA reads 12 KB of shared memory per 128 cycles); BN = 128 is used when
// the 256-wide grid would leave SMs idle (small diagonal blocks, few right-hand sides).
template <typename T, int BN> struct TcShape {
  static constexpr int A_BYTES = TC_BM * TcCfg<T>::BK * (int)sizeof(T);   // BM rows x BK elements of K (either majorness)
  static constexpr int B_BYTES = BN * TcCfg<T>::BK * (int)sizeof(T);
  static constexpr int HALF_STAGE = A_BYTES + B_BYTES;     // raw (= hi) tiles; the lo copies follow for Float32
  static constexpr int STAGE = HALF_STAGE * (TcCfg<T>::PASSES == 3 ? 2 : 1);
  static constexpr int STAGES = (sizeof(T) == 4 ? 2 : 1) * ((BN == 256) ? 2 : 3);   // Float16: 2 x 48 KB / 3 x 32 KB; Float32: 4 x 48 KB / 6 x 32 KB
  static constexpr int SMEM = STAGES * STAGE + 1024 + (TcCfg<T>::PASSES == 3 ? 16384 : 0);   // + drain staging for Float32
};

// 3xTF32 operand split.  kind::tf32 reads the top 19 bits of an FP32 word, i.e. it TRUNCATES the 13 low mantissa bits.
//   hi: the word itself (hi_round = 0: the tensor core's truncation is the split) or rounded to the nearest TF32 (hi_round = 0x1000,
//       stored masked); lo = a - hi is exact in FP32 either way.
//   lo has up to 13 significant bits of which the tensor core keeps 11: adding half a TF32 ulp to its bit pattern (lo_round = 0x1000)
//       turns that truncation into round-to-nearest -- without it every product is short by up to 2^-21 of its size, always in the
//       same direction, which is the dominant error of a long update chain (recursive LU in Float32: residual 5.9e-5 -> see DESIGN.md 4.4).
__device__ __forceinline__ void tf32_split(uint32_t v, uint32_t hi_round, uint32_t lo_round, uint32_t& h, uint32_t& l) {
  h = (v + hi_round) & 0xffffe000u;
  l = __float_as_uint(__uint_as_float(v) - __uint_as_float(h)) + lo_round;
}

struct GemmTcParams {
  int M, N, K;
  int a_mn0, a_k0;    // origin of the A operand in its tensor map (operand orientation: mn index, k index)
  int b_mn0, b_k0;
  void* C;            // top-left of the output block (column-major)
  long long ldc;
  float beta, sgn, post;
  int overwrite;      // 1: C <- post * sgn * A*B, the old contents of C are not read
  int raw_hi;         // Float32 (see tf32_split): 1 = leave the raw FP32 tile in place as the "hi" operand, 0 = hi rounded to nearest and stored, 2 = as 1 with lo truncated too.  kind::tf32 ignores the 13 low mantissa bits
                      // (measured on B200: results bit-identical to explicit masking), so only the lo tile has to be written.
  // K window per tile (batched TRMM, see nla_api.cu): t = tile index along M (or along N when win_on_n; N tile must be 128)
  //   win_mode 0: k in [0, K)                          1: diagonal block, k in [128t, 128t+128) of the V operand, [0,128) of the other
  //            2: k in [0, 128t)  (strictly below)     3: k in [128(t+1), K)  (strictly above)
  //            4: k in [0, 128(t+1))  (lower triangle with its diagonal block: block-inverse leaves, tri_inv.cuh)     5: k in [128t, K)  (upper)
  int win_mode, win_on_n, win_shift_a;
  unsigned long long* dbg;   // optional: 8 globaltimer stamps per CTA (probe builds; nullptr in production)
  int chunk_k;        // Float32: K extent accumulated in TMEM before it is promoted into C (0 = everything in one chunk)
  int tiles_m, tiles_n;
  // optional second destination for a sub-block of the final C (block-inverse leaves, nla_api.cu): C rows [dup_r0, +dup_rn) x
  // columns [dup_c0, +dup_cn) are also stored column-major at dup (leading dimension dup_ld)
  void* dup; long long dup_ld;
  int dup_r0, dup_rn, dup_c0, dup_cn;
  // conditioning guard of a block-inverse leaf (tri_guard.cuh): when the record says the block is ill-conditioned the launch does
  // nothing and the substitution kernel launched right behind it solves the block instead
  const double* skip_rec; double skip_thr2;
  // 0: whole block; 1 / 2: only the lower / upper triangle (diagonal included) of C is written (nla_lauum's diagonal blocks; one-tile kernel only)
  int tri_mode;
};

// evaluated AFTER griddepcontrol.wait (the record is written by a predecessor on the stream, or on a side stream joined by an event);
// uniform for the whole grid.  A skipped launch runs its prologue and teardown with zero K steps / zero tiles.
__device__ __forceinline__ bool tc_skip_launch(const GemmTcParams& p) {
  if (!p.skip_rec) return false;
  const double r0 = p.skip_rec[0], r1 = p.skip_rec[1], r2 = p.skip_rec[2];
  return r2 > 0.0 || !(r0 * r1 <= p.skip_thr2);
}

// instruction descriptor: FP32 accumulate, A/B format, majorness (0 = K-major, 1 = MN-major), N >> 3, M >> 4
template <typename T, int AMAJ, int BMAJ, int BN>
__host__ __device__ constexpr uint32_t tc_idesc() {
  return (1u << 4) | (TcCfg<T>::FMT << 7) | (TcCfg<T>::FMT << 10) | ((AMAJ == MAJ_MN ? 1u : 0u) << 15) | ((BMAJ == MAJ_MN ? 1u : 0u) << 16) |
         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

template <typename T> __device__ __forceinline__ T tc_from_float(float v);
template <> __device__ __forceinline__ float tc_from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __half tc_from_float<__half>(float v) { return __float2half_rn(v); }
template <typename T> __device__ __forceinline__ float tc_to_float(T v);
template <> __device__ __forceinline__ float tc_to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float tc_to_float<__half>(__half v) { return __half2float(v); }

template <typename T, int AMAJ, int BMAJ, int BN>
__global__ void __launch_bounds__(TcCfg<T>::THREADS, TcCfg<T>::MIN_CTAS)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const GemmTcParams p) {
  using Cfg = TcCfg<T>;
  using Shp = TcShape<T, BN>;
  constexpr int S = Shp::STAGES, BK = Cfg::BK, UK = Cfg::UK;
  constexpr int ES = (int)sizeof(T);
  constexpr int ATOM = 128 / ES;                       // elements of M/N in one 128-byte swizzle row (MN-major operands)
  constexpr int A_BYTES = Shp::A_BYTES, HALF_STAGE = Shp::HALF_STAGE, STAGE = Shp::STAGE;
  constexpr bool F32 = Cfg::PASSES == 3;
  constexpr int NBUF = Cfg::NBUF;                      // accumulator tiles in TMEM
  constexpr int TMEM_COLS = NBUF * BN;
  constexpr int DRAIN_WARP0 = Cfg::THREADS / 32 - 4;   // the last four warps own the four TMEM lane quarters

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[S];      // TMA bytes have landed
  __shared__ __align__(8) uint64_t conv_bar[S];      // Float32: lo tile written (128 arrivals)
  __shared__ __align__(8) uint64_t empty_bar[S];     // the MMAs that read this stage have completed
  __shared__ __align__(8) uint64_t dfull_bar[NBUF];  // a K chunk has been accumulated into this TMEM tile
  __shared__ __align__(8) uint64_t dfree_bar[NBUF];  // the drain warps have read it back (128 arrivals)
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  // grouped rasterisation: CTAs resident together share A row panels / B column panels in L2
  const int per_group = TC_GROUP_M * p.tiles_n;
  const int grp = blockIdx.x / per_group;
  const int first_m = grp * TC_GROUP_M;
  const int gsz = min(TC_GROUP_M, p.tiles_m - first_m);
  const int rem = blockIdx.x - grp * per_group;
  int tm = first_m + rem % gsz;
  int tn = rem / gsz;

  // K window of this tile
  int kbeg = 0, klen = p.K, ash = 0, bsh = 0;
  if (p.win_mode) {
    if (p.win_mode == 2 || p.win_mode == 4) {   // work grows with the tile index: start the heaviest tiles first
      if (p.win_on_n) tn = p.tiles_n - 1 - tn; else tm = p.tiles_m - 1 - tm;
    }
    const int t = p.win_on_n ? tn : tm;
    if (p.win_mode == 1) {
      klen = min(128, p.K - 128 * t);
      if (p.win_shift_a) ash = 128 * t; else bsh = 128 * t;
    } else if (p.win_mode == 2) {
      klen = min(p.K, 128 * t);
    } else if (p.win_mode == 4) {
      klen = min(p.K, 128 * (t + 1));
    } else if (p.win_mode == 5) {
      kbeg = 128 * t;
      klen = p.K - kbeg;
    } else {
      kbeg = 128 * (t + 1);
      klen = p.K - kbeg;
    }
    if (klen <= 0) return;   // uniform for the CTA; nothing allocated yet
  }

  // Programmatic dependent launch: let the next kernel of the stream start its own prologue (barrier init, TMEM allocation,
  // descriptor prefetch) while this one is still running; it blocks in griddepcontrol.wait until this grid has completed.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#define NLA_STAMP(slot) do { if (p.dbg) p.dbg[blockIdx.x * 8 + (slot)] = global_timer_ns(); } while (0)
  if (threadIdx.x == 0) NLA_STAMP(0);   // kernel entry

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; s++) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&conv_bar[s]), 128);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < NBUF; b++) {
      mbar_init(smem_u32(&dfull_bar[b]), 1);
      mbar_init(smem_u32(&dfree_bar[b]), 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), TMEM_COLS);
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&mapA); tma_prefetch_desc(&mapB); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // everything above touched no global data; from here on the previous kernel's results are needed
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) NLA_STAMP(1);   // prologue done (barriers, TMEM, dependency wait)
  const uint32_t tmem = tmem_slot;
  const int nk = tc_skip_launch(p) ? 0 : (klen + BK - 1) / BK;   // conditioning guard: a rejected block-inverse leaf does nothing
  // K is accumulated in TMEM in chunks; each finished chunk is added to C in registers with round-to-nearest by the drain
  // warps while the next chunk runs into the other TMEM tile.  The tensor core's own accumulation truncates, which biases
  // long same-sign sums (measured ~2^-24 relative per MMA); chunking bounds that for Float32.  Float16 uses one chunk.
  const int chunk_kb = (NBUF > 1 && p.chunk_k > 0) ? max(1, p.chunk_k / BK) : max(1, nk);
  const int nchunks = (nk + chunk_kb - 1) / chunk_kb;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const int am = p.a_mn0 + tm * TC_BM, bn = p.b_mn0 + tn * BN;
      for (int kt = 0; kt < nk; kt++) {
        const int s = kt % S, it = kt / S;
        if (it > 0) mbar_wait_wd(smem_u32(&empty_bar[s]), (it - 1) & 1);
        const uint32_t fb = smem_u32(&full_bar[s]);
        mbar_expect_tx(fb, HALF_STAGE);
        const uint32_t sa = smem_base + s * STAGE, sb = sa + A_BYTES;
        const int ak = p.a_k0 + ash + kbeg + kt * BK, bk = p.b_k0 + bsh + kbeg + kt * BK;
        if (AMAJ == MAJ_K) {
          tma_load_2d(sa, &mapA, fb, ak, am);                                            // box {BK, 128}
        } else {
#pragma unroll
          for (int a = 0; a < TC_BM / ATOM; a++) tma_load_2d(sa + a * (BK * 128), &mapA, fb, am + a * ATOM, ak);   // box {ATOM, BK}
        }
        if (BMAJ == MAJ_K) {
          tma_load_2d(sb, &mapB, fb, bk, bn);                                            // box {BK, BN}
        } else {
#pragma unroll
          for (int a = 0; a < BN / ATOM; a++) tma_load_2d(sb + a * (BK * 128), &mapB, fb, bn + a * ATOM, bk);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread drives the tensor core for the whole CTA =====
    if (lane == 0) {
      constexpr uint32_t idesc = tc_idesc<T, AMAJ, BMAJ, BN>();
      // K-major: 8-row groups 1024 B apart (SBO), K advances 32 B inside the 128-byte swizzle row.
      // MN-major: 8-k groups 1024 B apart (SBO), 128-byte M/N atoms BK*128 B apart (LBO), K advances UK rows of 128 B.
      constexpr uint32_t A_LBO = (AMAJ == MAJ_K) ? 16u : (uint32_t)(BK * 128), B_LBO = (BMAJ == MAJ_K) ? 16u : (uint32_t)(BK * 128);
      constexpr uint32_t A_KSTEP = (AMAJ == MAJ_K) ? 32u : (uint32_t)(UK * 128), B_KSTEP = (BMAJ == MAJ_K) ? 32u : (uint32_t)(UK * 128);
      // MN-major 32-bit operands: tcgen05 only accepts the 128-byte swizzle with 32-byte chunks, whose atom is 4 k-rows
      // (512 B) instead of 8 (the TMA map of such an operand uses CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B to match).
      constexpr bool A32 = (AMAJ == MAJ_MN) && ES == 4, B32 = (BMAJ == MAJ_MN) && ES == 4;
      // K-major operands: rows of BK*ES bytes (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B), 8-row groups 8 rows apart.
      constexpr uint32_t KROW = (uint32_t)(BK * ES);
      constexpr uint32_t K_LAY = (KROW == 128) ? UMMA_SW128 : UMMA_SW64;
      constexpr uint32_t A_SBO = (AMAJ == MAJ_K) ? 8u * KROW : (A32 ? 512u : 1024u), B_SBO = (BMAJ == MAJ_K) ? 8u * KROW : (B32 ? 512u : 1024u);
      constexpr uint32_t A_LAY = (AMAJ == MAJ_K) ? K_LAY : (A32 ? UMMA_SW128_BASE32B : UMMA_SW128);
      constexpr uint32_t B_LAY = (BMAJ == MAJ_K) ? K_LAY : (B32 ? UMMA_SW128_BASE32B : UMMA_SW128);
      static_assert(KROW == 128 || KROW == 64, "K-major rows must be 128 or 64 bytes");
      int kt = 0;
      bool stamped = false;
      for (int c = 0; c < nchunks; c++) {
        const int buf = c % NBUF, use = c / NBUF;
        if (use > 0) {   // the drain warps must have emptied this TMEM tile (chunk c - NBUF)
          mbar_wait_wd(smem_u32(&dfree_bar[buf]), (use - 1) & 1);
          tc_fence_after();
        }
        const uint32_t dt = tmem + (uint32_t)(buf * BN);
        uint32_t acc = 0;
        const int kend = min(nk, kt + chunk_kb);
        for (; kt < kend; kt++) {
          const int s = kt % S, it = kt / S;
          mbar_wait_wd(smem_u32(F32 ? &conv_bar[s] : &full_bar[s]), it & 1);
          tc_fence_after();
          if (!stamped) { NLA_STAMP(2); stamped = true; }   // first operand stage has landed
          const uint32_t sa = smem_base + s * STAGE, sb = sa + A_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / UK; kk++) {
            const uint64_t da = umma_desc(sa + kk * A_KSTEP, A_LBO, A_SBO, A_LAY);
            const uint64_t db = umma_desc(sb + kk * B_KSTEP, B_LBO, B_SBO, B_LAY);
            if (!F32) {
              tc_mma_f16(dt, da, db, idesc, acc);
            } else {
              const uint64_t dal = umma_desc(sa + HALF_STAGE + kk * A_KSTEP, A_LBO, A_SBO, A_LAY);
              const uint64_t dbl = umma_desc(sb + HALF_STAGE + kk * B_KSTEP, B_LBO, B_SBO, B_LAY);
              tc_mma_tf32(dt, dal, db, idesc, acc);   // lo * hi
              tc_mma_tf32(dt, da, dbl, idesc, 1u);    // hi * lo
              tc_mma_tf32(dt, da, db, idesc, 1u);     // hi * hi
            }
            acc = 1;
          }
          tc_commit(smem_u32(&empty_bar[s]));
        }
        tc_commit(smem_u32(&dfull_bar[buf]));
      }
      NLA_STAMP(3);   // all MMAs issued
    }
    __syncwarp();
  } else if (warp < DRAIN_WARP0) {
    // ===== Float32 only: split every landed tile into hi (the raw tile, see GemmTcParams::raw_hi) and lo =====
    const int et = threadIdx.x - 64;  // 0..127
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
      mbar_arrive(smem_u32(&conv_bar[s]));
    }
  } else {
    // ===== drain / epilogue: TMEM -> registers -> C.  A warp may only touch the TMEM lanes of its quarter (warp id mod 4). =====
    const int quarter = warp & 3;
    const int row = tm * TC_BM + quarter * 32 + lane;
    const bool row_ok = row < p.M;
    T* cbase = reinterpret_cast<T*>(p.C);
    T* crow = cbase + row;
    const int ncols = min(BN, p.N - tn * BN);   // valid columns of this tile
    // per-warp FP32 staging tile (32 columns x 32 rows).  Float16: the operand ring is idle once the single accumulator chunk is
    // complete, so its first 16 KB are reused; Float32 drains while the ring is live and has a dedicated region behind it.
    float* stg = reinterpret_cast<float*>(smem_gen + (F32 ? S * STAGE : 0)) + (warp - DRAIN_WARP0) * 1024;
    if (!p.overwrite) {
      // pull the old C tile into L2 while the main loop runs: each warp touches its own 32 rows of every column
      if (row_ok && (lane % (32 / ES)) == 0)   // one lane per 32-byte sector
        for (int j = 0; j < ncols; j++) asm volatile("prefetch.global.L2 [%0];" ::"l"(crow + (long long)(tn * BN + j) * p.ldc));
    }
#pragma unroll 1
    for (int c = 0; c < nchunks; c++) {
      const int buf = c % NBUF, use = c / NBUF;
      mbar_wait_wd(smem_u32(&dfull_bar[buf]), use & 1);
      tc_fence_after();
      const bool first = (c == 0), last = (c == nchunks - 1);
      if (last && threadIdx.x == DRAIN_WARP0 * 32) NLA_STAMP(4);   // last accumulator chunk complete
      const bool need_old = !(first && p.overwrite);
      // rounding points of the reference: B .= alpha .* B rounds to T (src/rectrxm.jl:64) -> beta on the first chunk; the update
      // rounds once per chunk (src/matmul.jl:64 rounds once per update); the trailing scale (src/rectrxm.jl:72) -> post on the last
      const float beta = first ? p.beta : 1.0f, post = last ? p.post : 1.0f;
      const uint32_t dt = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (c0 >= ncols) break;   // warp-uniform
        // TMEM -> registers (this thread: one row, 32 columns) -> per-warp FP32 staging tile [column][row] in shared memory
        uint32_t r[32];
        tmem_ld32(dt + (uint32_t)c0, r);
        tmem_ld_wait();
        __syncwarp();   // the previous pass has finished reading the staging tile
#pragma unroll
        for (int j = 0; j < 32; j++) stg[j * 32 + lane] = __uint_as_float(r[j]);
        __syncwarp();
        // shared memory -> C with 16-byte accesses: a lane owns VEC consecutive rows of one column, the warp covers CPP columns per pass
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
            } else if (col < ncols && grow < p.M) {   // ragged last rows: element-wise
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
            float accv[VEC];   // 16-byte shared-memory reads: consecutive lanes read consecutive 16/32-byte pieces (no bank conflicts)
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
            const int gcol = tn * BN + col;
            // triangle-masked store: rows [grow, grow + VEC) of column gcol against the diagonal of the output block
            const bool tri_all = p.tri_mode == 0 || (p.tri_mode == 1 ? grow >= gcol : grow + VEC - 1 <= gcol);
            const bool tri_none = p.tri_mode != 0 && (p.tri_mode == 1 ? grow + VEC - 1 < gcol : grow > gcol);
            if (tri_none) {
            } else if (grow + VEC <= p.M && tri_all) {
              *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(outv);
            } else {
#pragma unroll
              for (int e = 0; e < VEC; e++)
                if (grow + e < p.M && (p.tri_mode == 0 || (p.tri_mode == 1 ? grow + e >= gcol : grow + e <= gcol))) dst[e] = outv[e];
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
      if (NBUF > 1 && !last) {   // hand the TMEM tile back to the MMA issuer
        tc_fence_before();
        mbar_arrive(smem_u32(&dfree_bar[buf]));
      }
    }
  }

  // ===== teardown =====
  if (threadIdx.x == DRAIN_WARP0 * 32) NLA_STAMP(5);   // drain done
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
  if (threadIdx.x == 32) NLA_STAMP(6);   // TMEM released
#undef NLA_STAMP
}

}  // namespace nla

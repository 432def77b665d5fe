// gemm_tc.cuh -- Float16 / Float32 Schur/GEMM update on the 5th-generation tensor cores (tcgen05, accumulators in TMEM):
//     C <- post * (beta*C + sgn * opA(A) * opB(B))          (update)
//     C <- scale * opA(A) * opB(B)                          (overwrite: the diagonal-block leaves, see diag_prep.cuh)
//
// Replaces the reference's `matmul!` KernelAbstractions kernel (src/matmul.jl:5-66, launched from GEMM_ADD!/GEMM_SUB!
// :69-81; one output per work-item, accumulation in the element type) for the two low-precision element types.
//
// B200 design:
//  * CTA tile 128 x 256, one `tcgen05.mma.cta_group::1` of shape M128 x N256 x K(32 bytes) per step, issued by ONE thread;
//    the FP32 accumulator tile (128 lanes x 256 columns) lives in tensor memory.
//  * Float16: kind::f16 (FP16 inputs, FP32 accumulate -- north star; the reference accumulates in FP16).
//  * Float32: kind::tf32 three times per step ("3xTF32"): every operand tile is split in shared memory into
//    hi = a & 0xffffe000 (exactly a TF32 number) and lo = a - hi (exact in FP32), and D += hi*hi + hi*lo + lo*hi.  The
//    dropped lo*lo term and the TF32 rounding of lo leave a relative error of ~2^-21 per product, inside the reference's
//    1e-5 Float32 tolerance (test/trsm.jl:8); a single TF32 pass (2^-11) would not be.
//  * Operands come straight from the caller's column-major matrices through 2-D TMA tensor maps with SWIZZLE_128B:
//      K-major operand  (K contiguous in memory):  one box {128 B of K, 128|256 rows}
//      MN-major operand (M/N contiguous)        :  boxes {128 B of M/N, BK rows of K}, one per 128-byte M/N atom
//    and the tcgen05 shared-memory descriptors describe exactly those layouts (see umma_desc in common.cuh).  MN-major
//    Float32 operands use the 32-byte-chunk variant of the swizzle on both sides (the only one tcgen05 takes for them).
//  * Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = FP32 hi/lo
//    splitter during the main loop and epilogue (TMEM -> registers -> C) afterwards.  mbarrier ring between them;
//    `tcgen05.commit` releases a stage back to the producer and finally hands the accumulator to the epilogue.
//  * Float16 stages are 48 KB (2 stages, 2 CTAs per SM so one CTA's epilogue overlaps the other's main loop);
//    Float32 stages are 96 KB (raw/hi + lo copies, 2 stages, 1 CTA per SM).
#pragma once
#include "common.cuh"
#include "gemm_f64.cuh"  // MAJ_MN / MAJ_K

namespace nla {

constexpr int TC_BM = 128, TC_BN = 256;
constexpr int TC_THREADS = 192;
constexpr int TC_TMEM_COLS = 256;
constexpr int TC_GROUP_M = 8;

template <typename T> struct TcCfg;
template <> struct TcCfg<__half> {
  static constexpr int BK = 64, UK = 16, STAGES = 2, PASSES = 1, MIN_CTAS = 2;
  static constexpr uint32_t FMT = 0;  // F16
};
template <> struct TcCfg<float> {
  static constexpr int BK = 32, UK = 8, STAGES = 2, PASSES = 3, MIN_CTAS = 1;
  static constexpr uint32_t FMT = 2;  // TF32
};

template <typename T> __host__ __device__ constexpr int tc_a_bytes() { return TC_BM * 128; }   // BM rows x 128 B of K (or BK rows x BM elements)
template <typename T> __host__ __device__ constexpr int tc_b_bytes() { return TC_BN * 128; }
template <typename T> __host__ __device__ constexpr int tc_stage_bytes() { return (tc_a_bytes<T>() + tc_b_bytes<T>()) * (TcCfg<T>::PASSES == 3 ? 2 : 1); }
template <typename T> __host__ __device__ constexpr int tc_smem_bytes() { return TcCfg<T>::STAGES * tc_stage_bytes<T>() + 1024; }

struct GemmTcParams {
  int M, N, K;
  int a_mn0, a_k0;    // origin of the A operand in its tensor map (operand orientation: mn index, k index)
  int b_mn0, b_k0;
  void* C;            // top-left of the output block (column-major)
  long long ldc;
  float beta, sgn, post;
  int overwrite;      // 1: C <- (sgn*post) * A*B, the old contents of C are not read
  int tiles_m, tiles_n;
};

// instruction descriptor: FP32 accumulate, A/B format, majorness (0 = K-major, 1 = MN-major), N >> 3, M >> 4
template <typename T, int AMAJ, int BMAJ>
__host__ __device__ constexpr uint32_t tc_idesc() {
  return (1u << 4) | (TcCfg<T>::FMT << 7) | (TcCfg<T>::FMT << 10) | ((AMAJ == MAJ_MN ? 1u : 0u) << 15) | ((BMAJ == MAJ_MN ? 1u : 0u) << 16) |
         ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

template <typename T> __device__ __forceinline__ T tc_from_float(float v);
template <> __device__ __forceinline__ float tc_from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __half tc_from_float<__half>(float v) { return __float2half_rn(v); }
template <typename T> __device__ __forceinline__ float tc_to_float(T v);
template <> __device__ __forceinline__ float tc_to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float tc_to_float<__half>(__half v) { return __half2float(v); }

template <typename T, int AMAJ, int BMAJ>
__global__ void __launch_bounds__(TC_THREADS, TcCfg<T>::MIN_CTAS)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const GemmTcParams p) {
  using Cfg = TcCfg<T>;
  constexpr int S = Cfg::STAGES, BK = Cfg::BK, UK = Cfg::UK;
  constexpr int ES = (int)sizeof(T);
  constexpr int ATOM = 128 / ES;                       // elements of M/N in one 128-byte swizzle row (MN-major operands)
  constexpr int A_BYTES = tc_a_bytes<T>(), B_BYTES = tc_b_bytes<T>();
  constexpr int HALF_STAGE = A_BYTES + B_BYTES;        // raw (= hi) tiles; the lo copies follow for Float32
  constexpr int STAGE = tc_stage_bytes<T>();

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[S];    // TMA bytes have landed
  __shared__ __align__(8) uint64_t conv_bar[S];    // Float32: hi/lo split written (128 arrivals)
  __shared__ __align__(8) uint64_t empty_bar[S];   // the MMAs that read this stage have completed
  __shared__ __align__(8) uint64_t acc_bar;        // the whole accumulator tile is complete
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  // grouped rasterisation: CTAs resident together share A row panels / B column panels in L2
  const int per_group = TC_GROUP_M * p.tiles_n;
  const int grp = blockIdx.x / per_group;
  const int first_m = grp * TC_GROUP_M;
  const int gsz = min(TC_GROUP_M, p.tiles_m - first_m);
  const int rem = blockIdx.x - grp * per_group;
  const int tm = first_m + rem % gsz;
  const int tn = rem / gsz;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; s++) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&conv_bar[s]), 128);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&acc_bar), 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int nk = (p.K + BK - 1) / BK;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      tma_prefetch_desc(&mapA);
      tma_prefetch_desc(&mapB);
      const int am = p.a_mn0 + tm * TC_BM, bn = p.b_mn0 + tn * TC_BN;
      for (int kt = 0; kt < nk; kt++) {
        const int s = kt % S, it = kt / S;
        if (it > 0) mbar_wait_wd(smem_u32(&empty_bar[s]), (it - 1) & 1);
        const uint32_t fb = smem_u32(&full_bar[s]);
        mbar_expect_tx(fb, HALF_STAGE);
        const uint32_t sa = smem_base + s * STAGE, sb = sa + A_BYTES;
        const int ak = p.a_k0 + kt * BK, bk = p.b_k0 + kt * BK;
        if (AMAJ == MAJ_K) {
          tma_load_2d(sa, &mapA, fb, ak, am);                                            // box {BK, 128}
        } else {
#pragma unroll
          for (int a = 0; a < TC_BM / ATOM; a++) tma_load_2d(sa + a * (BK * 128), &mapA, fb, am + a * ATOM, ak);   // box {ATOM, BK}
        }
        if (BMAJ == MAJ_K) {
          tma_load_2d(sb, &mapB, fb, bk, bn);                                            // box {BK, 256}
        } else {
#pragma unroll
          for (int a = 0; a < TC_BN / ATOM; a++) tma_load_2d(sb + a * (BK * 128), &mapB, fb, bn + a * ATOM, bk);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread drives the tensor core for the whole CTA =====
    if (lane == 0) {
      constexpr uint32_t idesc = tc_idesc<T, AMAJ, BMAJ>();
      // K-major: 8-row groups 1024 B apart (SBO), K advances 32 B inside the 128-byte swizzle row.
      // MN-major: 8-k groups 1024 B apart (SBO), 128-byte M/N atoms BK*128 B apart (LBO), K advances UK rows of 128 B.
      constexpr uint32_t A_LBO = (AMAJ == MAJ_K) ? 16u : (uint32_t)(BK * 128), B_LBO = (BMAJ == MAJ_K) ? 16u : (uint32_t)(BK * 128);
      constexpr uint32_t A_KSTEP = (AMAJ == MAJ_K) ? 32u : (uint32_t)(UK * 128), B_KSTEP = (BMAJ == MAJ_K) ? 32u : (uint32_t)(UK * 128);
      // MN-major 32-bit operands: tcgen05 only accepts the 128-byte swizzle with 32-byte chunks, whose atom is 4 k-rows
      // (512 B) instead of 8 (the TMA map of such an operand uses CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B to match).
      constexpr bool A32 = (AMAJ == MAJ_MN) && ES == 4, B32 = (BMAJ == MAJ_MN) && ES == 4;
      constexpr uint32_t A_SBO = A32 ? 512u : 1024u, B_SBO = B32 ? 512u : 1024u;
      constexpr uint32_t A_LAY = A32 ? UMMA_SW128_BASE32B : UMMA_SW128, B_LAY = B32 ? UMMA_SW128_BASE32B : UMMA_SW128;
      uint32_t acc = 0;
      for (int kt = 0; kt < nk; kt++) {
        const int s = kt % S, it = kt / S;
        mbar_wait_wd(smem_u32(Cfg::PASSES == 3 ? &conv_bar[s] : &full_bar[s]), it & 1);
        tc_fence_after();
        const uint32_t sa = smem_base + s * STAGE, sb = sa + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / UK; kk++) {
          const uint64_t da = umma_desc(sa + kk * A_KSTEP, A_LBO, A_SBO, A_LAY);
          const uint64_t db = umma_desc(sb + kk * B_KSTEP, B_LBO, B_SBO, B_LAY);
          if (Cfg::PASSES == 1) {
            tc_mma_f16(tmem, da, db, idesc, acc);
            acc = 1;
          } else {
            const uint64_t dal = umma_desc(sa + HALF_STAGE + kk * A_KSTEP, A_LBO, A_SBO, A_LAY);
            const uint64_t dbl = umma_desc(sb + HALF_STAGE + kk * B_KSTEP, B_LBO, B_SBO, B_LAY);
            tc_mma_tf32(tmem, dal, db, idesc, acc);   // lo * hi
            tc_mma_tf32(tmem, da, dbl, idesc, 1u);    // hi * lo
            tc_mma_tf32(tmem, da, db, idesc, 1u);     // hi * hi
            acc = 1;
          }
        }
        tc_commit(smem_u32(&empty_bar[s]));
      }
      tc_commit(smem_u32(&acc_bar));
    }
    __syncwarp();
  } else {
    const int et = threadIdx.x - 64;  // 0..127
    if (Cfg::PASSES == 3) {
      // ===== Float32: split every landed tile into hi (in place) and lo =====
      for (int kt = 0; kt < nk; kt++) {
        const int s = kt % S, it = kt / S;
        mbar_wait_wd(smem_u32(&full_bar[s]), it & 1);
        uint4* hi = reinterpret_cast<uint4*>(smem_gen + s * STAGE);
        uint4* lo = reinterpret_cast<uint4*>(smem_gen + s * STAGE + HALF_STAGE);
#pragma unroll 4
        for (int i = et; i < HALF_STAGE / 16; i += 128) {
          uint4 v = hi[i], h, l;
          h.x = v.x & 0xffffe000u; h.y = v.y & 0xffffe000u; h.z = v.z & 0xffffe000u; h.w = v.w & 0xffffe000u;
          l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
          l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
          l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
          l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
          hi[i] = h;
          lo[i] = l;
        }
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&conv_bar[s]));
      }
    }
    // ===== epilogue: TMEM -> registers -> C.  A warp may only touch the TMEM lanes of its quarter (warp id mod 4). =====
    mbar_wait_wd(smem_u32(&acc_bar), 0);
    tc_fence_after();
    const int quarter = warp & 3;
    const int row = tm * TC_BM + quarter * 32 + lane;
    const bool row_ok = row < p.M;
    T* crow = reinterpret_cast<T*>(p.C) + row;
    const float scale = p.sgn * p.post;
    const bool plain = (p.beta == 1.0f) && (p.post == 1.0f);
#pragma unroll 1
    for (int c0 = 0; c0 < TC_BN; c0 += 32) {
      const int colbase = tn * TC_BN + c0;
      if (colbase >= p.N) break;   // warp-uniform
      uint32_t r[32];
      tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
      float old[32];
      if (!p.overwrite) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
          old[j] = 0.f;
          if (row_ok && colbase + j < p.N) old[j] = tc_to_float<T>(crow[(long long)(colbase + j) * p.ldc]);
        }
      }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j++) {
        if (row_ok && colbase + j < p.N) {
          const float a = __uint_as_float(r[j]);
          float v;
          if (p.overwrite) {
            v = scale * a;
          } else if (plain) {
            v = old[j] + p.sgn * a;
          } else {
            // rounding points of the reference: B .= alpha .* B rounds to T (src/rectrxm.jl:64), the update rounds once
            // (src/matmul.jl:64), the trailing scale rounds again (src/rectrxm.jl:72)
            v = tc_to_float<T>(tc_from_float<T>(p.beta * old[j])) + p.sgn * a;
            if (p.post != 1.0f) v = p.post * tc_to_float<T>(tc_from_float<T>(v));
          }
          crow[(long long)(colbase + j) * p.ldc] = tc_from_float<T>(v);
        }
      }
    }
  }

  // ===== teardown =====
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TC_TMEM_COLS);
}

}  // namespace nla

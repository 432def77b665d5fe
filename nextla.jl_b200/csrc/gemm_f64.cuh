// gemm_f64.cuh -- FP64 Schur/GEMM update on the tensor cores:  C <- post * (beta*C + sgn * opA(A) * opB(B)).
//
// Replaces the reference's `matmul!` KernelAbstractions kernel (src/matmul.jl:5-66, launched from
// GEMM_ADD!/GEMM_SUB! :69-81), which computes one output per work-item from two 32x32 shared tiles.
//
// B200 design (see DESIGN.md "FP64 GEMM"):
//  * tcgen05 has no f64 kind on sm_100a, so FP64 tensor math is warp-level `mma.sync.m8n8k4.f64`
//    (SASS DMMA.8x8x4, measured peak 37.0 TFLOP/s = 16 cycles per DMMA per SM sub-partition).
//  * CTA tile 128x128x16, 8 consumer warps (2 x 4, each 64x32 -> 32 accumulator fragments) + 1 producer warp.
//  * Operands are staged by TMA (cp.async.bulk.tensor.3d) into a 4-stage mbarrier ring.  Both operand
//    layouts use 64-byte rows with SWIZZLE_64B: a 3-D tensor map {8 elements, outer dim, blocks-of-8}
//    turns a column-major panel into [block][outer][8] rows, and the lane->k assignment
//        k(q, step) = (q&1) + 4*(q>>1) + 2*(step&1) + 8*(step>>1)
//    makes every 64-bit fragment load bank-conflict free for both the MN-major and the K-major layout
//    (the sum over k is order independent, so permuting k inside a 16-wide K tile is legal as long as
//    the A and B fragments use the same permutation).
//  * Out-of-range K (only ever at the matrix boundary, see schedule.h) is zero-filled by TMA for both
//    operands; out-of-range M/N rows/columns produce accumulators that the epilogue never stores.
#pragma once
#include "common.cuh"

namespace nla {

constexpr int GF_BM = 128, GF_BN = 128, GF_BK = 16, GF_STAGES = 4;
constexpr int GF_TILE_BYTES = GF_BM * GF_BK * 8;          // 16 KB per operand tile
constexpr int GF_STAGE_BYTES = 2 * GF_TILE_BYTES;         // 32 KB
constexpr int GF_CONSUMER_WARPS = 8;
constexpr int GF_THREADS = (GF_CONSUMER_WARPS + 1) * 32;  // + 1 TMA producer warp
constexpr int GF_SMEM_BYTES = GF_STAGES * GF_STAGE_BYTES + 1024;  // + alignment slack
constexpr int GF_GROUP_M = 16;                            // rasterisation: 16 M-tiles x all N-tiles per group

enum { MAJ_MN = 0, MAJ_K = 1 };

struct GemmF64Params {
  int M, N, K;       // extents of this update
  int a_mn0, a_k0;   // origin of the A operand inside its parent matrix (elements, operand orientation)
  int b_mn0, b_k0;   // origin of the B operand inside its parent matrix
  double* C;         // top-left of the output block (column-major)
  long long ldc;
  double beta, sgn, post;
  int tiles_m, tiles_n;
  int tri_mode;    // 0: whole block; 1 / 2: only the lower / upper triangle (diagonal included) of C is read and written (nla_lauum)
  int overwrite;   // 1: C <- post * sgn * A*B (old contents not read)
};

// byte offset of element (outer index `o`, inner index 0..7 `e`) inside one 64B-swizzled block-of-8 panel:
// row pitch 64 B, 16-byte chunk index XORed with bits [1:2] of the outer index (CU_TENSOR_MAP_SWIZZLE_64B).
__device__ __forceinline__ uint32_t sw64(uint32_t o, uint32_t e) { return o * 64u + ((((e >> 1) ^ ((o >> 1) & 3u)) << 4) | ((e & 1u) << 3)); }

template <int AMAJ, int BMAJ>
__global__ void __launch_bounds__(GF_THREADS, 1)
gemm_f64_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const GemmF64Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[GF_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[GF_STAGES];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

  // grouped rasterisation so that the CTAs resident together share A row-panels and B column-panels in L2
  const int per_group = GF_GROUP_M * p.tiles_n;
  const int grp = blockIdx.x / per_group;
  const int first_m = grp * GF_GROUP_M;
  const int gsz = min(GF_GROUP_M, p.tiles_m - first_m);
  const int rem = blockIdx.x - grp * per_group;
  const int tm = first_m + rem % gsz;
  const int tn = rem / gsz;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < GF_STAGES; s++) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), GF_CONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int nk = (p.K + GF_BK - 1) / GF_BK;

  if (warp == GF_CONSUMER_WARPS) {
    // ===== TMA producer: one elected lane =====
    if (lane == 0) {
      tma_prefetch_desc(&mapA);
      tma_prefetch_desc(&mapB);
      const int am = p.a_mn0 + tm * GF_BM, bn = p.b_mn0 + tn * GF_BN;
      for (int kt = 0; kt < nk; kt++) {
        const int s = kt % GF_STAGES, it = kt / GF_STAGES;
        if (it > 0) mbar_wait(smem_u32(&empty_bar[s]), (it - 1) & 1);
        const uint32_t fb = smem_u32(&full_bar[s]);
        mbar_expect_tx(fb, GF_STAGE_BYTES);
        const uint32_t sa = smem_base + s * GF_STAGE_BYTES, sb = sa + GF_TILE_BYTES;
        const int ak = p.a_k0 + kt * GF_BK, bk = p.b_k0 + kt * GF_BK;
        if (AMAJ == MAJ_MN) tma_load_3d(sa, &mapA, fb, 0, ak, am >> 3);   // box {8, 16, 16}: [mblock][k][8]
        else                tma_load_3d(sa, &mapA, fb, 0, am, ak >> 3);   // box {8, 128, 2}: [kblock][m][8]
        if (BMAJ == MAJ_K)  tma_load_3d(sb, &mapB, fb, 0, bn, bk >> 3);
        else                tma_load_3d(sb, &mapB, fb, 0, bk, bn >> 3);
      }
    }
    return;
  }

  // ===== consumers: 8 warps, warp tile 64 (M) x 32 (N) =====
  const int wm = warp >> 2, wn = warp & 3;
  const uint32_t g = lane >> 2, q = lane & 3;

  uint32_t aoff[4], boff[4];
#pragma unroll
  for (int st = 0; st < 4; st++) {
    const uint32_t s1 = st & 1, s2 = st >> 1;
    const uint32_t ki = (q & 1) + 4 * (q >> 1) + 2 * s1;  // k inside its block of 8
    const uint32_t k = ki + 8 * s2;                       // k inside the 16-wide tile
    if (AMAJ == MAJ_MN) aoff[st] = (wm * 8) * 1024u + sw64(k, g);                       // + i*1024
    else                aoff[st] = s2 * (GF_BM * 64u) + sw64(wm * 64 + g, ki);          // + i*512
    if (BMAJ == MAJ_K)  boff[st] = GF_TILE_BYTES + s2 * (GF_BN * 64u) + sw64(wn * 32 + g, ki);  // + j*512
    else                boff[st] = GF_TILE_BYTES + (wn * 4) * 1024u + sw64(k, g);              // + j*1024
  }
  constexpr uint32_t ASTR = (AMAJ == MAJ_MN) ? 1024u : 512u;
  constexpr uint32_t BSTR = (BMAJ == MAJ_K) ? 512u : 1024u;

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int kt = 0; kt < nk; kt++) {
    const int s = kt % GF_STAGES;
    mbar_wait(smem_u32(&full_bar[s]), (kt / GF_STAGES) & 1);
    const uint32_t sbase = smem_base + s * GF_STAGE_BYTES;
#pragma unroll
    for (int st = 0; st < 4; st++) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a[i]) : "r"(sbase + aoff[st] + i * ASTR));
#pragma unroll
      for (int j = 0; j < 4; j++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(b[j]) : "r"(sbase + boff[st] + j * BSTR));
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&empty_bar[s]));
  }

  // ===== epilogue: C <- post * (beta*C + sgn*acc), straight from the accumulator fragments =====
  const int row0 = tm * GF_BM + wm * 64 + g;
  const int col0 = tn * GF_BN + wn * 32 + 2 * q;
  const bool unit = (p.beta == 1.0) && (p.post == 1.0);
#pragma unroll
  for (int j = 0; j < 4; j++) {
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const int col = col0 + j * 8 + c;
      if (col < p.N) {
        double* cp = p.C + (long long)col * p.ldc;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int row = row0 + i * 8;
          if (row < p.M && !(p.tri_mode == 1 ? row < col : (p.tri_mode == 2 && row > col))) {
            if (p.overwrite) { cp[row] = __dmul_rn(p.post, p.sgn * acc[i][j][c]); continue; }
            double v = cp[row];
            if (unit) {
              v = v + p.sgn * acc[i][j][c];
            } else {
              v = __dmul_rn(p.beta, v);
              v = __dadd_rn(v, p.sgn * acc[i][j][c]);
              v = __dmul_rn(p.post, v);
            }
            cp[row] = v;
          }
        }
      }
    }
  }
}

}  // namespace nla

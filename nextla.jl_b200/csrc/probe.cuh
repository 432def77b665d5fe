// probe.cuh -- FP64 tensor-core issue-rate microbenchmark (nla_probe_fp64_peak): the measured denominator of the FP64 roofline.
// MEASURED_PEAKS.json carries HBM and bf16 figures only, so bench.py measures the DMMA.8x8x4 peak of the very GPU it runs on: every
// warp of a full SM issues back-to-back independent mma.sync.m8n8k4.f64 (2 x 2 accumulator fragments per warp, 32 warps per SM -- the
// configuration that reached 37.04 TFLOP/s = 16.08 cycles per DMMA per sub-partition in profiles/r01_probe_dmma_peak.txt).
#pragma once
#include "common.cuh"

namespace nla {

__global__ void __launch_bounds__(1024) dmma_peak_kernel(double* out, int iters) {
  double a[2], b[2], c[8];
#pragma unroll
  for (int i = 0; i < 2; i++) { a[i] = threadIdx.x * 1e-9 + i; b[i] = threadIdx.x * 1e-9 - i; }
#pragma unroll
  for (int i = 0; i < 8; i++) c[i] = 0.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 2; j++) dmma884(c[(i * 2 + j) * 2], c[(i * 2 + j) * 2 + 1], a[i], b[j]);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i];
  if (s == 123.456) out[0] = s;   // never true: keeps the loop alive
}

}  // namespace nla

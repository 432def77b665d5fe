/* CPU restatement (C + OpenMP) of NextLA.jl's unified_rectrxm! path.  TEST INFRASTRUCTURE ONLY:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * Same algorithm as oracle/reference_port.py (which is checked against it in tests/test_oracle.py);
 * this build exists so that sizes beyond a few hundred finish in seconds and so that bench.py can time
 * "the reference's CPU path" (work-groups spread over host threads) on the GPU box's cores.
 * Parity status: see the header of oracle/reference_port.py ("pinned by the reference's own test
 * criterion"; no golden vectors exist upstream; Julia is not installed so the reference cannot run). */
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define T double
#define SFX f64
#include "nla_oracle_impl.h"
#undef T
#undef SFX

#define T float
#define SFX f32
#include "nla_oracle_impl.h"
#undef T
#undef SFX

#define T _Float16
#define SFX f16
#include "nla_oracle_impl.h"
#undef T
#undef SFX

/* dtype: 0 = Float64, 1 = Float32, 2 = Float16 (same enum as include/nextla_b200.h) */
int nla_oracle_rectrxm(char side, char uplo, char trans, char func, int dtype, int64_t n, int64_t m, double alpha,
                       void *A, int64_t lda, void *B, int64_t ldb) {
  if (n <= 0 || m <= 0) return 0;
  switch (dtype) {
    case 0: rectrxm_f64(side, uplo, trans, func, n, m, alpha, (double *)A, lda, (double *)B, ldb); return 0;
    case 1: rectrxm_f32(side, uplo, trans, func, n, m, alpha, (float *)A, lda, (float *)B, ldb); return 0;
    case 2: rectrxm_f16(side, uplo, trans, func, n, m, alpha, (_Float16 *)A, lda, (_Float16 *)B, ldb); return 0;
  }
  return 1;
}

int nla_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

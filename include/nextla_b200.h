/* nextla_b200.h -- C ABI of the B200-native replacement for NextLA.jl's recursive TRSM/TRMM path.
 *
 * Every entry point replaces one piece of the reference's Julia interface (paths relative to the
 * NextLA.jl repository).  All matrices are column-major (Julia layout), device pointers unless the name
 * says `_host`, element type selected by `dtype`.  All calls are asynchronous on `stream` (the reference
 * does not synchronise either, src/rectrxm.jl:75) and return an int status (0 = NLA_OK); nothing throws
 * or aborts across the ABI.  There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef NEXTLA_B200_H
#define NEXTLA_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nla_context *nla_handle_t;

enum nla_dtype {
  NLA_F64 = 0, NLA_F32 = 1, NLA_F16 = 2, /* Float64 / Float32 / Float16 (src/rectrxm.jl:101: T<:AbstractFloat) */
  NLA_C64 = 3, NLA_C128 = 4              /* ComplexF32 / ComplexF64, interleaved (re, im): nla_rectrxm_complex only */
};

enum nla_status {
  NLA_OK = 0,
  NLA_ERR_INVALID_CHAR = 1,   /* side/uplo/trans/diag/func not in the documented set (the reference silently coerces, src/rectrxm.jl:105-123) */
  NLA_ERR_INVALID_DIM = 2,    /* negative n/m or leading dimension too small */
  NLA_ERR_INVALID_DTYPE = 3,
  NLA_ERR_NULL_POINTER = 4,
  NLA_ERR_CUDA = 5,           /* a CUDA runtime/driver call failed; nla_last_cuda_error() has the code */
  NLA_ERR_NO_DEVICE = 6,
  NLA_ERR_UNSUPPORTED = 7,
  NLA_ERR_INVALID_HANDLE = 8,
  NLA_ERR_WORKSPACE = 9,      /* caller-provided workspace (nla_set_workspace) too small even for the degraded schedule */
  NLA_ERR_NCCL = 10           /* libnccl could not be loaded, or an NCCL call failed (nla_mg_*) */
};

/* Library lifetime.  One handle per (host thread / task, device, stream); re-entrant across handles, no process-global mutable state
 * (per-kernel attributes, helper streams and workspaces all live in the handle).  A handle owns device workspaces (prepared diagonal
 * blocks / block inverses of the Float32/Float16 path, a copy of B for the batched multiply and the Float64 right side): calls through
 * ONE handle must be ordered on ONE stream -- use one handle per concurrently used stream (the Julia and Python bindings key their
 * handle cache by (device, stream)).  Every entry point runs on the handle's device and restores the caller's current device. */
int nla_create(nla_handle_t *handle, int device);
int nla_destroy(nla_handle_t handle);
const char *nla_status_string(int status);
int nla_last_cuda_error(nla_handle_t handle);
int nla_version(void);

/* Workspace control.  By default the handle's workspaces grow on demand with the STREAM-ORDERED allocator
 * (cudaMallocAsync on the call's stream): no device synchronisation, the call stays asynchronous even when it has to grow one.
 *   nla_workspace_bytes  bytes the call (side, func, dtype, n, m) needs under the handle's current options (0 for Float64 left side)
 *   nla_reserve          pre-size the library-owned workspaces (and create the helper streams) for that call: afterwards calls of that
 *                        shape or smaller allocate nothing
 *   nla_set_workspace    use a caller-owned device arena (256-byte aligned) instead: the library then never allocates on the
 *                        nla_rectrxm / nla_trxm / nla_gemm_update path.  If the arena is smaller than nla_workspace_bytes the call degrades
 *                        (128-wide leaves instead of block inverses, in-place instead of batched multiply, native instead of transposed
 *                        right side) and only returns NLA_ERR_WORKSPACE when even that does not fit.  (NULL, 0) returns to library-owned. */
int64_t nla_workspace_bytes(nla_handle_t handle, char side, char func, int dtype, int64_t n, int64_t m);
int nla_reserve(nla_handle_t handle, char side, char func, int dtype, int64_t n, int64_t m);
int nla_set_workspace(nla_handle_t handle, void *workspace, int64_t bytes);

/* unified_rectrxm!(side, uplo, transpose, alpha, func, A, B)            -- src/rectrxm.jl:43-76 (+ unified_rec :101-198)
 *   func 'S': B <- alpha * op(A)^-1 * B (side 'L')  or  alpha * B * op(A)^-1 (side 'R')
 *   func 'M': B <- alpha * op(A) * B                or  alpha * B * op(A)
 * A is n x n (only the `uplo` triangle is read, non-unit diagonal), B is n x m (side 'L') or m x n (side 'R'),
 * m = number of independent right-hand-side vectors.  'C' == 'T' for the real element types supported. */
int nla_rectrxm(nla_handle_t handle, char side, char uplo, char trans, char func, int dtype, int64_t n, int64_t m,
                double alpha, const void *A, int64_t lda, void *B, int64_t ldb, void *stream);

/* BLAS-style entry with the `diag` flag: trsm(side, uplo, transa, diag, A, B, alpha) / trmm(...)   -- src/trsm.jl:186-205, src/trmm.jl:430-448
 * (SURVEY.md 8(f1)).  The reference's wrappers accept `transa` and `diag` and ignore both (and do not recurse: n <= 1024 / 16); here
 * `trans` is honoured, the recursion is the one of nla_rectrxm, and diag = 'U' treats the diagonal of A as ones without reading it
 * (what the unit-lower solve of a recursive LU needs, src/lu.jl:277).  diag = 'N' is exactly nla_rectrxm. */
int nla_trxm(nla_handle_t handle, char side, char uplo, char trans, char diag, char func, int dtype, int64_t n, int64_t m,
             double alpha, const void *A, int64_t lda, void *B, int64_t ldb, void *stream);

/* Complex element types: the same operation for ComplexF32 / ComplexF64 matrices (interleaved re/im, Julia layout),
 * complex alpha, BLAS `diag` flag, and trans = 'C' (conjugate transpose) DISTINCT from 'T'.  The reference advertises complex support
 * (README.md:20) and builds Adjoint(A) for 'C' (src/rectrxm.jl:57) but restricts the recursion to real types (:101), so a complex call
 * fails there.  Runs the reference's recursion on planar copies of A and B in a library workspace (4 n^2 + 4 n m reals): every complex
 * update is four real GEMM updates on the tensor-core kernels of the real path, diagonal blocks by complex substitution. */
int nla_rectrxm_complex(nla_handle_t handle, char side, char uplo, char trans, char diag, char func, int dtype, int64_t n, int64_t m,
                        double alpha_re, double alpha_im, const void *A, int64_t lda, void *B, int64_t ldb, void *stream);

/* Same operation with HOST buffers (pinned or pageable): stages A once and streams B through the device in
 * RHS slabs, overlapping copies with compute; synchronous (returns when B_host holds the result).
 * This is the end-to-end path bench.py reports as `e2e`. */
int nla_rectrxm_host(nla_handle_t handle, char side, char uplo, char trans, char func, int dtype, int64_t n, int64_t m,
                     double alpha, const void *A_host, int64_t lda, void *B_host, int64_t ldb);

/* Multi-GPU: the right-hand sides are sharded across GPUs and A is broadcast from its owner.  So that the
 * broadcast overlaps the solve, A may arrive in panels of `panel_cols` columns: panel p = columns [p*panel_cols, (p+1)*panel_cols),
 * and panel_events[p] is a cudaEvent_t recorded (on any stream of this device) after the copy / ncclBroadcast that fills it.
 * The schedule makes `stream` wait for a panel's event right before the first launch that reads it.  (The Float32/Float16 path
 * prepares the whole diagonal first and therefore waits for every panel up front.)  The reference has no multi-device path.
 * nla_panel_order: host-only, the order in which the schedule first touches the panels = the order to broadcast them in;
 * writes up to max_panels indices and returns the number of panels (or a negative nla_status). */
int nla_rectrxm_gated(nla_handle_t handle, char side, char uplo, char trans, char func, int dtype, int64_t n, int64_t m,
                      double alpha, const void *A, int64_t lda, void *B, int64_t ldb, void *stream, int64_t panel_cols,
                      int64_t n_panels, void *const *panel_events);
int64_t nla_panel_order(char side, char uplo, char trans, char func, int64_t n, int64_t panel_cols, int64_t *order,
                        int64_t max_panels);

/* Single-process multi-GPU: one host thread drives every GPU of the box; the library owns the NCCL communicators
 * (ncclCommInitAll; libnccl is bound at run time with dlopen, so a process that already carries NCCL shares its copy), per-GPU compute /
 * broadcast / upload streams, the per-panel events and the replicas of A.  The right-hand sides are sharded by the caller: shard i
 * (shard_m[i] vectors, leading dimension ldb[i]) lives on GPU i -- columns of B for side 'L', rows for side 'R'.
 *   nla_mg_create(&mg, ngpu, devices)   devices = NULL: GPUs 0 .. ngpu-1.  ngpu = 1 needs no NCCL.
 *   nla_mg_rectrxm      A on GPU `root` (device pointer), B shards on their GPUs.  A is broadcast in column panels in the order the schedule
 *                       consumes them, every GPU's solve gated panel by panel (nla_rectrxm_gated).  ASYNCHRONOUS on the library's streams:
 *                       A and the shards must be complete when the call is made; nla_mg_sync waits for the results.
 *   nla_mg_rectrxm_host A and the B shards in HOST memory (pinned for full speed).  The k-th consumed panel of A is uploaded through GPU
 *                       (k mod ngpu)'s own PCIe link -- only the referenced trapezoid -- and broadcast from there; every GPU streams its shard
 *                       of B through the host pipeline (nla_rectrxm_hostb_gated), one host thread per GPU inside the call.  Synchronous.
 *   nla_mg_handle(mg, i) the per-GPU handle (for nla_set_option); nla_mg_stream(mg, i) GPU i's compute stream (cudaStream_t). */
typedef struct nla_mg_context *nla_mg_t;
int nla_mg_create(nla_mg_t *mg, int ngpu, const int *devices);
int nla_mg_destroy(nla_mg_t mg);
int nla_mg_device_count(nla_mg_t mg);
nla_handle_t nla_mg_handle(nla_mg_t mg, int i);
void *nla_mg_stream(nla_mg_t mg, int i);
int nla_mg_last_nccl_error(nla_mg_t mg);
int nla_mg_sync(nla_mg_t mg);
int nla_mg_rectrxm(nla_mg_t mg, char side, char uplo, char trans, char func, int dtype, int64_t n, double alpha, int root,
                   const void *A_root, int64_t lda, void *const *B_shards, const int64_t *shard_m, const int64_t *ldb);
int nla_mg_rectrxm_host(nla_mg_t mg, char side, char uplo, char trans, char func, int dtype, int64_t n, double alpha,
                        const void *A_host, int64_t lda, void *const *B_host_shards, const int64_t *shard_m, const int64_t *ldb);

/* Host B, device A: the multi-GPU end-to-end path.  A is (becoming) resident on this device -- panel_events as in nla_rectrxm_gated,
 * n_panels = 0 if it is already complete -- while this rank's right-hand sides live in host memory: B_host is streamed through
 * the device exactly as in nla_rectrxm_host (chunks in first-touch order, copied back as soon as they are final).  Synchronous. */
int nla_rectrxm_hostb_gated(nla_handle_t handle, char side, char uplo, char trans, char func, int dtype, int64_t n, int64_t m,
                            double alpha, const void *A_dev, int64_t lda, void *B_host, int64_t ldb, int64_t panel_cols,
                            int64_t n_panels, void *const *panel_events);

/* Strided (pitched) asynchronous copy on `stream`: to_device = 1 host -> device, 0 device -> host, 2 device -> device.  The multi-GPU
 * driver uses it to upload, and to pack / unpack for the broadcast, only the referenced triangle of a column panel of A (a trapezoid:
 * `height` columns of `width_bytes` each). */
int nla_memcpy2d_async(nla_handle_t handle, void *dst, int64_t dst_pitch_bytes, const void *src, int64_t src_pitch_bytes,
                       int64_t width_bytes, int64_t height, int to_device, void *stream);

/* laswp(A, first, last, ipiv, incx)                                                   -- src/lu.jl:470-530
 * Row interchanges on a device matrix (rows x ncols, column-major): for i = k1..k2 (incx = 1) or k2..k1 (incx = -1) swap rows i and
 * ipiv[i] (1-based, like the reference; ipiv is a DEVICE vector of int64 = Julia Int, entries outside [k1, k2] are not read).  With
 * nla_trxm(..., diag = 'U') and nla_gemm_update this runs the laswp + TRSM + GEMM steps of the reference's recursive LU
 * (getrf2!, src/lu.jl:274-280, :297) on the device; nla_getrf2 below is the whole factorisation.
 * A forward walk is planned once per call (which row ends where, in batches of 16 pivots) and applied with independent gathers and
 * scatters; a backward walk goes pivot by pivot. */
int nla_laswp(nla_handle_t handle, int dtype, int64_t rows, int64_t ncols, void *A, int64_t lda, int64_t k1, int64_t k2,
              const int64_t *ipiv, int incx, void *stream);

/* getrf2!(A, ipiv, info)                                                                  -- src/lu.jl:185-299
 * Recursive LU with partial pivoting of a device matrix, A = P * L * U, every step on the device: the reference's recursion on the
 * columns (split at min(m, n) / 2, :253-256) with laswp (:274, :298), the unit-lower recursive TRSM (:277) and the GEMM update (:280) of
 * this library, and a panel kernel (getrf.cuh) for blocks of <= 64 columns -- one thread-block cluster exchanging pivot candidates through
 * distributed shared memory when the panel's rows fit it, a cooperative grid-wide launch otherwise -- that applies the reference's
 * single-column rule (:224-251): pivot = first entry of largest magnitude, interchange, scale by the reciprocal (divide when
 * |pivot| < sfmin), a zero pivot leaves the column alone and is reported.
 *   A     m x n, column-major, leading dimension lda >= max(1, m); overwritten by L (unit diagonal not stored) and U
 *   ipiv  DEVICE vector of min(m, n) int64, 1-based: row i was interchanged with row ipiv[i] (the reference's Vector{Int})
 *   info  DEVICE int: 0, or i if U[i, i] is exactly zero (first such i); the reference returns it, a device-side caller reads it after
 *         synchronising the stream.  Asynchronous: no host synchronisation inside.
 * NLA_F64 and NLA_F32 (the reference's real BlasFloat types); NLA_F16 and the complex types return NLA_ERR_UNSUPPORTED.
 * Illegal sizes (m < 0, n < 0, lda < max(1, m); the reference's info = -1 / -2 / -4, :192-203) return NLA_ERR_INVALID_DIM. */
int nla_getrf2(nla_handle_t handle, int dtype, int64_t m, int64_t n, void *A, int64_t lda, int64_t *ipiv, int *info, void *stream);

/* lauum!(uplo, n, A, ib)                                                                 -- src/lauum.jl:52-186
 * A := L^H * L (uplo 'L') or U * U^H (uplo 'U') on a device matrix: the triangular factor is stored in the `uplo` triangle of A, the
 * result replaces it in the same triangle; the opposite triangle is neither used nor written (the reference multiplies whole diagonal
 * blocks, i.e. assumes it holds zeros).  Block loop of the reference (compute_lower! / compute_upper!) with the off-diagonal steps on the
 * trmm / GEMM kernels of this library and the diagonal-block products through GEMMs whose epilogue stores only the `uplo` triangle.
 * ib <= 0: 1024.  The reference throws ArgumentError for a bad uplo or negative n (:54-60): NLA_ERR_INVALID_CHAR / NLA_ERR_INVALID_DIM. */
int nla_lauum(nla_handle_t handle, char uplo, int dtype, int64_t n, void *A, int64_t lda, int64_t ib, void *stream);

/* Diagonal-block leaves: LeftLowerTRSM!/LeftUpperTRSM!/RightLowerTRSM!/RightUpperTRSM!  -- src/trsm.jl:128-150
 * and LeftLowerTRMM!/.../RightUpperTRMM!                                               -- src/trmm.jl:332-389.
 * n <= nla_leaf_max(dtype) = 1024, the reference's cap (its kernels hold the block in 1024-entry shared arrays, src/trsm.jl:9-11):
 * one launch of the leaf kernel up to n = 128, the library's blocked path (fused slab / block inverses) above. */
int nla_trsm_leaf(nla_handle_t handle, char side, char uplo, int dtype, int64_t n, int64_t m,
                  const void *A, int64_t lda, void *B, int64_t ldb, void *stream);
int nla_trmm_leaf(nla_handle_t handle, char side, char uplo, int dtype, int64_t n, int64_t m,
                  const void *A, int64_t lda, void *B, int64_t ldb, void *stream);
int64_t nla_leaf_max(int dtype);

/* GEMM_ADD!(A,B,C): C += A*B and GEMM_SUB!(A,B,C): A -= B*C                            -- src/matmul.jl:69-81
 * exposed as one update: C(MxN) <- C + sign * opA(A)(MxK) * opB(B)(KxN), sign = +1 or -1, opX = 'N' or 'T'
 * (the reference passes transposition through Transpose wrappers, src/matmul.jl:30-32,40-42). */
int nla_gemm_update(nla_handle_t handle, int dtype, char transa, char transb, int64_t M, int64_t N, int64_t K, int sign,
                    const void *A, int64_t lda, const void *B, int64_t ldb, void *C, int64_t ldc, void *stream);

/* Tunables (the reference hard-codes its thresholds, src/rectrxm.jl:52,63).  Keys:
 *   "leaf"        recursion cutoff = diagonal-block size handled by one leaf launch (default per dtype)
 *   "force_simt"  1 = never use the tensor-core GEMM kernels (debug / A-B comparison)
 *   "macro"       order of the diagonal blocks solved by the fused slab kernel (FP64 left side): -1 (default) = automatic -- the whole
 *                 diagonal (ONE launch for the call) when the call has at least 48 x #SM right-hand sides, 2048 otherwise, at most one
 *                 panel in a gated call; 0 = off; any other value = that block order
 *   "host_stream" 1 (default) = a Float64 left-side solve from host buffers (nla_rectrxm_host) with at least 48 x #SM right-hand sides runs
 *                 as one streaming launch of the row-split slab kernel: operands arrive chunk by chunk behind device flags, finished
 *                 chunks are downloaded while the kernel runs; 0 = the chunked multi-launch pipeline
 *   "gated_stream" 0 (default); 1 = nla_rectrxm_gated runs a Float64 left-side solve with at least 48 x #SM right-hand sides as one launch too,
 *                 every block row waiting on a device word that counts the arrived panels.  ONLY safe when the panels are delivered by
 *                 copy engines / peer DMA: the kernel occupies every SM while it waits, so SM-based collectives (NCCL kernels) queued
 *                 behind it never start
 *   "gated_macro" fused-slab block order of a gated call (nla_rectrxm_gated, Float64 left side): default 2048 (measured per step, 2048 vs 4096: 130.5 vs 129.4 ms at 2 GPUs, 130.1-130.7 vs 131.7 ms at 8 GPUs); never less than one panel
 *   "host_macro", "host_macro_mid"  chunked host pipeline: fused-slab block order at both ends / in the middle of the diagonal (1024 / 1024)
 *   "slab_kind"   fused FP64 slab kernel: 0 (default) = row-split (the 8 consumer warps share the 128 rows of a block row; CTA width 112 or
 *                 56 vectors, whichever fills the 148 SMs best for the call's number of right-hand sides), 1 = column-split (128 / 64 vectors)
 *   "slab_w"      right-hand-side vectors per CTA of the fused slab kernel: 0 = automatic; 112 / 56 force a width of the row-split kernel,
 *                 128 / 64 select the column-split kernel with that width
 *   "streams"     number of RHS slabs run on concurrent streams (0 = automatic: one per 4096 vectors, at most 4)
 *   "tc_bn"       N tile of the Float32/Float16 tcgen05 GEMM: 0 = automatic (256, or 128 when the 256-wide grid would not fill the SMs), 128, 256
 *   "tc_cg"       CTA pairs (tcgen05 cta_group::2, 256 x 256 tile per pair) for the large updates: 0 = automatic, 1 = never, 2 = whenever M > 128
 *   "tf32_raw_hi" Float32 3xTF32 split: 1 (default) = the raw FP32 tile is the hi operand (the tensor core drops the low 13 bits), 0 = mask explicitly
 *   "tc_chunk_k"  Float32: K extent accumulated in tensor memory before it is added into C with round-to-nearest (default 512; 0 = never)
 *   "trmm_batched" Float32/Float16 multiply: 1 (default) = out-of-place batched schedule (one copy of B in the handle's workspace, all diagonal
 *                 blocks in one launch, one triangular GEMM), 0 = the reference's in-place recursion
 *   "inv_block"   Float32/Float16 solve: order of the diagonal blocks that are inverted once per call (FP64 128-blocks doubled in the next
 *                 wider type) so that a leaf is ONE triangular tcgen05 GEMM: 0 = default (1024), 128 = plain 128-wide leaves, powers of two up to 4096
 *   "right_via_left" Float64, side 'R': 1 (default) = run the equivalent left-side problem (same Teff) on a transposed copy of B in the handle's
 *                 workspace (n*m*8 bytes, at most 8 GiB) so that the fused slab kernel applies; 0 = the native right-side schedule
 *   "host_slabs"  host-buffer entry points, Float64: number of independent RHS slabs, each on its own compute stream with its own part of every
 *                 chunk of B (0 = automatic: one per 4096 vectors, at most 4; 1 = one stream)
 *   "tc_wide_k"   Float16: updates with K >= this value run on 256 x 512 pair tiles (persistent, two accumulators share the A tile: a third
 *                 less L2 traffic per flop); default 4096, 0 = never
 *   "tc_persist"  Float16: 1 (default) = the persistent CTA-pair kernel (static tile list per cluster, two TMEM accumulators, 8 drain warps)
 *                 takes every multi-tile launch with K < 8192 and the block-inverse leaves; 0 = never; 2 = also the long updates
 *   "inv_dup"     1 (default) = an update also writes the block of B the next block-inverse leaf reads into the leaf's workspace (0: one copy per leaf)
 *   "inv_overlap" 1 (default) = all but the first two block inverses are computed on a side stream while the solve is running
 *   "pdl"         1 (default) = launch the tcgen05 kernels with programmatic dependent launch (prologue overlaps the predecessor's tail)
 *   "inv_guard"   Float32/Float16 solve with block inverses: 1 (default) = conditioning guard.  Every inverted block gets a device-side
 *                 record (||T||_F, ||inv T||_F, non-finite entries); a block whose ||T||_F ||inv T||_F / order exceeds 1/(32 eps) (64 for
 *                 Float16, 524288 for Float32) or whose rounded inverse is not finite is solved by SUBSTITUTION (the reference's leaf
 *                 arithmetic, src/trsm.jl:15-27; backward stable for any conditioning) instead of the inverse GEMM.  Decided on the device,
 *                 the call stays asynchronous.  0 = off: the solve is then only conditionally stable (error ~ eps * cond(block)).
 *   "inv_guard_kappa" threshold override of the guard (0 = the default above; 1 rejects every block = substitution leaves everywhere)
 *   "nvtx"        1 = NVTX ranges around every call and schedule op (nsys / ncu --nvtx)
 *   "getrf_cluster" nla_getrf2 panels: -1 (default) = one thread-block cluster of 16 CTAs (8 if the device refuses 16) whenever the
 *                 panel's rows fit its shared memory, else the grid-wide cooperative kernel; 0 = grid-wide kernel only; 8 / 16 = that size
 *   "profile"     1 = bracket every kernel launch with CUDA events (read back with nla_profile_read)
 * Read-only keys of nla_get_option:
 *     "inv_fallbacks" (synchronises) number of blocks of the last guarded solve that took the substitution fallback
 *     "ws_allocs"     device allocations made by the library for this handle so far  */
int nla_set_option(nla_handle_t handle, const char *key, int64_t value);
int64_t nla_get_option(nla_handle_t handle, const char *key);

/* Host-only introspection of the schedule that replaces the recursive splitter (src/rectrxm.jl:101-198): writes up to
 * max_ops records of 6 int64 {kind (0 = leaf, 1 = GEMM update), c0, cn, k0, kn, carries_alpha} in launch order, in the
 * normalised coordinates of DESIGN.md (leaf: diagonal block [c0, c0+cn); update: V[c0:c0+cn] +-= Teff[c,k] V[k0:k0+kn]).
 * `leaf` = recursion cutoff: <= 0 -> 128; up to 4096 (4096 = the fused FP64 slab schedule, 1024 = the block-inverse Float32/Float16 one).
 * Returns the number of ops (may exceed max_ops) or a negative nla_status.  Needs no GPU. */
int64_t nla_plan(char side, char uplo, char trans, char func, int64_t n, int64_t leaf, int64_t *ops, int64_t max_ops);

/* Host-only introspection of the host-buffer pipeline (nla_rectrxm_host / nla_rectrxm_hostb_gated): the schedule with recursion cutoff
 * `cutoff`, its large updates cut into 1024-wide pieces, for `slabs` RHS slabs, and the transfer plan that feeds it.  Writes up to
 * max_ops op records (6 int64 as in nla_plan), up to max_xfers transfer records of 3 int64 {kind (0 = 1024 x 1024 tile (i, j) of A, 1 = chunk i
 * of 1024 vector elements of B for slab j), i, j} in issue order, *n_xfers = number of transfers, need_out[op * slabs + slab] = index of the
 * last transfer that op needs for that slab (-1 = none), last_out[chunk] = index of the last op that writes the chunk (its download
 * follows that op).  a_resident = 1: A is not staged (nla_rectrxm_hostb_gated).  Returns the number of ops or a negative nla_status. */
int64_t nla_host_plan(char side, char uplo, char trans, char func, int64_t n, int64_t cutoff, int64_t slabs, int a_resident,
                      int64_t *ops_out, int64_t max_ops, int64_t *xfers_out, int64_t max_xfers, int64_t *n_xfers, int64_t *need_out,
                      int64_t *last_out);

/* Per-launch device times recorded while option "profile" is 1: waits for the recorded work, writes up to max_records
 * records of 3 doubles {kind (0 = leaf, 1 = GEMM update), algorithmic flops of the launch, milliseconds} in launch order,
 * clears the log and returns the number of records that were available (or a negative nla_status).  If there is room, one more record
 * of kind 2 follows: the device time from the start of the first launch to the end of the last (launches + the gaps between them). */
int64_t nla_profile_read(nla_handle_t handle, double *records, int64_t max_records);

/* Measured FP64 tensor-core peak of the handle's GPU: runs a DMMA.8x8x4 issue-rate microbenchmark (a few milliseconds, synchronous)
 * and writes TFLOP/s.  bench.py's FP64 roofline denominator (MEASURED_PEAKS.json has no FP64 figure). */
int nla_probe_fp64_peak(nla_handle_t handle, double *tflops);

/* Counters for bench.py: kernels launched by this handle since the last reset. */
int64_t nla_launch_count(nla_handle_t handle, int reset);

#ifdef __cplusplus
}
#endif
#endif /* NEXTLA_B200_H */

/* Included three times by nla_oracle.c with T / ACC_SUFFIX defined.  TEST INFRASTRUCTURE ONLY.
 * CPU restatement of NextLA.jl's unified_rectrxm! path; every function cites the reference
 * file:line it follows (paths relative to /root/reference).  A "view" is (ptr, row stride, col stride)
 * so that Transpose/SubArray wrappers (src/rectrxm.jl:57,137-148) stay zero-copy as in Julia. */

#define V(v, i, j) ((v).p[(int64_t)(i) * (v).rs + (int64_t)(j) * (v).cs])

typedef struct { T *p; int64_t rs, cs; } CAT(view_, SFX);

static CAT(view_, SFX) CAT(sub_, SFX)(CAT(view_, SFX) v, int64_t i0, int64_t j0) {
  CAT(view_, SFX) r = { v.p + i0 * v.rs + j0 * v.cs, v.rs, v.cs };
  return r;
}

/* matmul! kernel, src/matmul.jl:5-66: one work-group per 32x32 output tile, one output per work-item;
 * per K-tile a partial sum `tmp` in T over the zero-padded tile (:50-53), `outval += tmp` in T (:54),
 * finally output += alpha*outval evaluated in Float64 (alpha::Float64, :6,:64), rounded once to T.
 * Work-groups are spread over OpenMP threads the way KernelAbstractions' CPU backend spreads them. */
static void CAT(matmul_, SFX)(CAT(view_, SFX) out, CAT(view_, SFX) in1, CAT(view_, SFX) in2,
                             int64_t N, int64_t R, int64_t M, double alpha) {
  const int TD = 32;
  int64_t gN = (N + TD - 1) / TD, gM = (M + TD - 1) / TD, nt = (R + TD - 1) / TD;
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t gj = 0; gj < gM; gj++)
    for (int64_t gi = 0; gi < gN; gi++) {
      T tile1[32][33], tile2[32][33], outval[32][32];
      for (int i = 0; i < TD; i++) for (int j = 0; j < TD; j++) outval[i][j] = (T)0;
      for (int64_t t = 0; t < nt; t++) {
        for (int j = 0; j < TD; j++)
          for (int i = 0; i < TD; i++) {
            int64_t I = gi * TD + i, J = gj * TD + j, K1 = t * TD + j, K2 = t * TD + i;
            tile1[i][j] = (I < N && K1 < R) ? V(in1, I, K1) : (T)0;   /* :28-35 */
            tile2[i][j] = (K2 < R && J < M) ? V(in2, K2, J) : (T)0;   /* :37-45 */
          }
        for (int j = 0; j < TD; j++)
          for (int i = 0; i < TD; i++) {
            int64_t I = gi * TD + i, J = gj * TD + j;
            if (I < N && J < M) {
              T tmp = (T)0;
              for (int k = 0; k < TD; k++) tmp = (T)(tmp + (T)(tile1[i][k] * tile2[k][j]));  /* :50-53 */
              outval[i][j] = (T)(outval[i][j] + tmp);                                      /* :54 */
            }
          }
      }
      for (int j = 0; j < TD; j++)
        for (int i = 0; i < TD; i++) {
          int64_t I = gi * TD + i, J = gj * TD + j;
          if (I < N && J < M) V(out, I, J) = (T)((double)V(out, I, J) + alpha * (double)outval[i][j]);  /* :64 */
        }
    }
}

/* TRSM leaves, src/trsm.jl:5-126.  One work-group per RHS vector.  `left`: vectors are columns of B and
 * the system is Teff x = b; right: vectors are rows of B and x Teff = b.  `fwd` selects pivot order
 * 1..n (lower_left_kernel :21-27, right_upper_kernel :114-120) or n..1 (upper_left :52-58, right_lower :83-89).
 * x_r <- b_r/d_r (:15-18), then x_r -= (a/d_r)*x_i for every r "after" pivot i, with a = A[r,i] (left) or
 * A[i,r] (right) of the *parent* view (the launchers' Transpose(A) at :131,:149 cancels the kernels' swapped index). */
static void CAT(trsm_leaf_, SFX)(int left, int fwd, CAT(view_, SFX) A, CAT(view_, SFX) B, int64_t n, int64_t m) {
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < m; v++) {
    T diag[1024], x[1024];
    for (int64_t r = 0; r < n; r++) {
      diag[r] = V(A, r, r);
      T b = left ? V(B, r, v) : V(B, v, r);
      x[r] = (T)(b / diag[r]);
    }
    if (fwd) {
      for (int64_t i = 0; i < n; i++)
        for (int64_t r = i + 1; r < n; r++) {
          T a = left ? V(A, r, i) : V(A, i, r);
          T s = (T)(a / diag[r]);
          x[r] = (T)(x[r] - (T)(s * x[i]));
        }
    } else {
      for (int64_t i = n - 1; i >= 1; i--)
        for (int64_t r = 0; r < i; r++) {
          T a = left ? V(A, r, i) : V(A, i, r);
          T s = (T)(a / diag[r]);
          x[r] = (T)(x[r] - (T)(s * x[i]));
        }
    }
    for (int64_t r = 0; r < n; r++) { if (left) V(B, r, v) = x[r]; else V(B, v, r) = x[r]; }
  }
}

/* TRMM leaves, src/trmm.jl:43-312 (n <= 16): out = sum over the triangular k range, sequential in T.
 * left/lower k=1..i (:91-93); left/upper k=i..N (:164-166); right/lower k=j..N (:233-235); right/upper k=1..j (:296-298). */
static void CAT(trmm_leaf_, SFX)(int left, int lower, CAT(view_, SFX) A, CAT(view_, SFX) B, int64_t n, int64_t m) {
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < m; v++) {
    T b[16], o[16];
    for (int64_t r = 0; r < n; r++) b[r] = left ? V(B, r, v) : V(B, v, r);
    for (int64_t r = 0; r < n; r++) {
      T acc = (T)0;
      if (left) {
        int64_t k0 = lower ? 0 : r, k1 = lower ? r : n - 1;
        for (int64_t k = k0; k <= k1; k++) acc = (T)(acc + (T)(V(A, r, k) * b[k]));
      } else {
        int64_t k0 = lower ? r : 0, k1 = lower ? n - 1 : r;
        for (int64_t k = k0; k <= k1; k++) acc = (T)(acc + (T)(b[k] * V(A, k, r)));
      }
      o[r] = acc;
    }
    for (int64_t r = 0; r < n; r++) { if (left) V(B, r, v) = o[r]; else V(B, v, r) = o[r]; }
  }
}

/* unified_rec, src/rectrxm.jl:101-198. */
static void CAT(rec_, SFX)(int solve, int left, int lower, CAT(view_, SFX) A, int64_t n, CAT(view_, SFX) B,
                          int64_t m, int64_t threshold) {
  if (n <= threshold) {                                   /* :103-126 */
    if (solve) CAT(trsm_leaf_, SFX)(left, left ? lower : !lower, A, B, n, m);
    else       CAT(trmm_leaf_, SFX)(left, lower, A, B, n, m);
    return;
  }
  int64_t mid;                                            /* :129-134 */
  if ((n & (n - 1)) == 0) mid = n / 2; else { mid = 1; while (mid * 2 < n) mid *= 2; }
  int64_t rem = n - mid;
  CAT(view_, SFX) A11 = A, A22 = CAT(sub_, SFX)(A, mid, mid), A21 = CAT(sub_, SFX)(A, mid, 0), A12 = CAT(sub_, SFX)(A, 0, mid);
  CAT(view_, SFX) B1 = B, B2 = left ? CAT(sub_, SFX)(B, mid, 0) : CAT(sub_, SFX)(B, 0, mid);   /* :143-149 */
  int forward = (left && lower && solve) || (!left && !lower && solve) || (left && !lower && !solve) || (!left && lower && !solve); /* :153-156 */
  if (forward) {
    CAT(rec_, SFX)(solve, left, lower, A11, mid, B1, m, threshold);
    if (left) {
      if (solve) CAT(matmul_, SFX)(B2, A21, B1, rem, mid, m, -1.0);   /* :164 B2 -= A21*B1 */
      else       CAT(matmul_, SFX)(B1, A12, B2, mid, rem, m, 1.0);    /* :166 B1 += A12*B2 */
    } else {
      if (solve) CAT(matmul_, SFX)(B2, B1, A12, m, mid, rem, -1.0);   /* :170 B2 -= B1*A12 */
      else       CAT(matmul_, SFX)(B1, B2, A21, m, rem, mid, 1.0);    /* :172 B1 += B2*A21 */
    }
    CAT(rec_, SFX)(solve, left, lower, A22, rem, B2, m, threshold);
  } else {
    CAT(rec_, SFX)(solve, left, lower, A22, rem, B2, m, threshold);
    if (left) {
      if (solve) CAT(matmul_, SFX)(B1, A12, B2, mid, rem, m, -1.0);   /* :184 B1 -= A12*B2 */
      else       CAT(matmul_, SFX)(B2, A21, B1, rem, mid, m, 1.0);    /* :186 B2 += A21*B1 */
    } else {
      if (solve) CAT(matmul_, SFX)(B1, B2, A21, m, rem, mid, -1.0);   /* :190 B1 -= B2*A21 */
      else       CAT(matmul_, SFX)(B2, B1, A12, m, mid, rem, 1.0);    /* :192 B2 += B1*A12 */
    }
    CAT(rec_, SFX)(solve, left, lower, A11, mid, B1, m, threshold);
  }
}

/* unified_rectrxm!, src/rectrxm.jl:43-76.  `B .= alpha .* B` is evaluated in Float64 (alpha is a Float64 in the
 * reference's tests) and rounded to T; before the recursion for 'S' (:62-65), after it for 'M' (:71-73). */
static void CAT(rectrxm_, SFX)(char side, char uplo, char trans, char func, int64_t n, int64_t m, double alpha,
                              T *Ap, int64_t lda, T *Bp, int64_t ldb) {
  CAT(view_, SFX) A = { Ap, 1, lda }, B = { Bp, 1, ldb };
  int left = side == 'L', lower = uplo == 'L', solve = func == 'S';
  int64_t threshold = 16;
  if (trans == 'T' || trans == 'C') { A.rs = lda; A.cs = 1; lower = !lower; }    /* :56-59 */
  int64_t br = left ? n : m, bc = left ? m : n;
  if (solve) {
    threshold = 256;
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < bc; j++) for (int64_t i = 0; i < br; i++) V(B, i, j) = (T)(alpha * (double)V(B, i, j));
  }
  CAT(rec_, SFX)(solve, left, lower, A, n, B, m, threshold);
  if (!solve) {
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < bc; j++) for (int64_t i = 0; i < br; i++) V(B, i, j) = (T)(alpha * (double)V(B, i, j));
  }
}
#undef V

/* oracle/oracle_body.h -- TEST INFRASTRUCTURE (CPU oracle), included twice by oracle.c with
 *   REAL = double / float, FMA = fma / fmaf, SQRT = sqrt / sqrtf, NAME(x) = x##_f64 / x##_f32.
 *
 * A plain-C restatement of the reference's sparse iterative-solve hot path.  Each function cites the
 * reference lines it follows (paths relative to /root/reference).  It reproduces not only the algorithm but the
 * ROUNDING of the reference as built with g++ -O3 -march=x86-64-v4|v3 (oracle/Makefile):
 *   - GCC contracts `acc += a*b` into one fused multiply-add (-ffp-contract=fast is GCC's default), including
 *     across Eigen's packet intrinsics, so every a*b+c below is an explicit FMA() -- except the SpMV row loop;
 *   - Eigen's reductions keep TWO packet accumulators of `lanes` lanes each and fold them with a fixed
 *     horizontal-add pattern (Core/Redux.h:227-282 and the predux<> of each ISA), reproduced by NAME(redux).
 * `lanes` = 8/4/2 doubles (16/8/4 floats) selects the AVX-512 / AVX / SSE pattern; lanes = 1 is a plain
 * left-to-right sum.  With the matching `lanes` the restatement is bit-identical to oracle/_ref (checked by
 * tests/test_oracle_pinned.py); the CUDA path is compared against it with the tolerances of the north star.
 */

/* ---- y = A*x, CSR, one right-hand side --------------------------------------------------------------------
 * SparseCore/SparseDenseProduct.h:26-72 (row-major, ColPerCol): for each row `tmp=0; tmp += v*x[c]` left to
 * right, then `res(i) += alpha*tmp` on a destination zeroed by Core/ProductEvaluators.h:348-349 (alpha == 1).
 * ROUNDING: in this one loop g++ 13.3 -O3 does NOT contract (it emits an in-order reduction of separately
 * rounded products, for both ISA levels; established by probing oracle/_ref with random rows of length 1..79),
 * so the product and the add are two roundings here -- unlike every other a*b+c of the path. */
void NAME(oracle_spmv_blk)(int64_t rows, const int32_t* rowptr, const int32_t* colidx, const REAL* vals,
                           const REAL* x, REAL* y, int blk) {
  /* blk == 0: every product rounded separately (what g++ emits for double).  blk == V > 0: the first
   * floor(len/V)*V entries of a row are unfused (the vectorised in-order body) and the remaining < V entries are
   * FMAs (the scalar epilogue) -- what g++ emits for float, V = 8 at x86-64-v4 and 4 at x86-64-v3 (probed). */
  for (int64_t i = 0; i < rows; ++i) {
    REAL tmp = (REAL)0;
    int32_t k = rowptr[i], end = rowptr[i + 1];
    int32_t body_end = blk > 0 ? k + ((end - k) / blk) * blk : end;
    for (; k < body_end; ++k) tmp = tmp + vals[k] * x[colidx[k]];
    for (; k < end; ++k) tmp = FMA(vals[k], x[colidx[k]], tmp);
    y[i] = (REAL)0 + tmp; /* 0 + 1*tmp: turns -0 into +0 exactly as the reference's += does */
  }
}

void NAME(oracle_spmv)(int64_t rows, const int32_t* rowptr, const int32_t* colidx, const REAL* vals,
                       const REAL* x, REAL* y) {
  NAME(oracle_spmv_blk)(rows, rowptr, colidx, vals, x, y, 0);
}

/* ---- y = selfadjointView<UpLo>(A)*x from ONE stored triangle of a row-major matrix ---------------------------
 * SparseCore/SparseSelfAdjointView.h:279-337.  Entries of the other triangle, if stored, are ignored.
 * uplo: 1 = Lower (ProcessFirstHalf for row-major), 2 = Upper (ProcessSecondHalf). */
void NAME(oracle_symv)(int64_t n, const int32_t* rowptr, const int32_t* colidx, const REAL* vals, const REAL* x,
                       REAL* y, int uplo) {
  for (int64_t i = 0; i < n; ++i) y[i] = (REAL)0;
  for (int64_t j = 0; j < n; ++j) {
    int32_t k = rowptr[j], end = rowptr[j + 1];
    if (uplo == 2) { /* :307-315 skip the strictly-lower part, take the diagonal first */
      while (k < end && colidx[k] < j) ++k;
      if (k < end && colidx[k] == j) {
        y[j] = FMA(vals[k], x[j], y[j]); /* alpha*value*rhs, alpha==1; contracted into the += */
        ++k;
      }
    }
    REAL xj = x[j]; /* :318 alpha*rhs(j) */
    REAL acc = (REAL)0;
    for (; (uplo == 1) ? (k < end && colidx[k] < j) : (k < end); ++k) { /* :321-327 gather + scatter */
      REAL a = vals[k];
      acc = FMA(a, x[colidx[k]], acc);
      y[colidx[k]] = FMA(a, xj, y[colidx[k]]);
    }
    y[j] += acc;                                                         /* :328 */
    if (uplo == 1 && k < end && colidx[k] == j) y[j] = FMA(vals[k], x[j], y[j]); /* :331-332 */
  }
}

/* ---- Jacobi preconditioner setup ------------------------------------------------------------------------------
 * IterativeLinearSolvers/BasicPreconditioners.h:64-79: first stored entry with inner index == j; missing or
 * exactly zero diagonal -> 1. */
void NAME(oracle_jacobi_factorize)(int64_t n, const int32_t* rowptr, const int32_t* colidx, const REAL* vals,
                                   REAL* invdiag) {
  for (int64_t j = 0; j < n; ++j) {
    int32_t k = rowptr[j], end = rowptr[j + 1];
    while (k < end && colidx[k] != j) ++k;
    invdiag[j] = (k < end && vals[k] != (REAL)0) ? (REAL)1 / vals[k] : (REAL)1;
  }
}

/* ---- sum_i a[i]*b[i] with the reference's accumulation order ---------------------------------------------------
 * Core/Dot.h:67-99 -> Core/Redux.h:227-282 (LinearVectorizedTraversal, NoUnrolling) on a cwiseProduct / abs2
 * expression (alignedStart == 0): two packet accumulators striding 2*lanes, folded acc0+acc1, one odd packet,
 * predux (arch/AVX512/PacketMath.h:982-988, arch/AVX/PacketMath.h:647-650, arch/SSE/PacketMath.h:722-731 and
 * their float twins: every float predux is a halving tree), then the scalar tail. */
static REAL NAME(predux)(const REAL* p, int lanes) {
  REAL t[16];
  for (int l = 0; l < lanes; ++l) t[l] = p[l];
  if (sizeof(REAL) == 8 && lanes == 8) { /* hadd pattern of predux<Packet8d> */
    REAL s0 = t[0] + t[4], s1 = t[1] + t[5], s2 = t[2] + t[6], s3 = t[3] + t[7];
    return (s0 + s1) + (s2 + s3);
  }
  for (int h = lanes / 2; h >= 1; h /= 2) /* upper half added onto lower half */
    for (int l = 0; l < h; ++l) t[l] = t[l] + t[l + h];
  return t[0];
}

REAL NAME(oracle_dot)(const REAL* a, const REAL* b, int64_t n, int lanes) {
  if (lanes <= 1) {
    if (n == 0) return (REAL)0;
    REAL res = a[0] * b[0];
    for (int64_t i = 1; i < n; ++i) res = FMA(a[i], b[i], res);
    return res;
  }
  const int64_t P = lanes;
  const int64_t size2 = (n / (2 * P)) * (2 * P), size1 = (n / P) * P;
  REAL res;
  if (size1) {
    REAL acc0[16], acc1[16];
    for (int l = 0; l < P; ++l) acc0[l] = a[l] * b[l];
    if (size1 > P) {
      for (int l = 0; l < P; ++l) acc1[l] = a[P + l] * b[P + l];
      for (int64_t i = 2 * P; i < size2; i += 2 * P)
        for (int l = 0; l < P; ++l) {
          acc0[l] = FMA(a[i + l], b[i + l], acc0[l]);
          acc1[l] = FMA(a[i + P + l], b[i + P + l], acc1[l]);
        }
      for (int l = 0; l < P; ++l) acc0[l] = acc0[l] + acc1[l];
      if (size1 > size2)
        for (int l = 0; l < P; ++l) acc0[l] = FMA(a[size2 + l], b[size2 + l], acc0[l]);
    }
    res = NAME(predux)(acc0, (int)P);
    for (int64_t i = size1; i < n; ++i) res = FMA(a[i], b[i], res);
  } else {
    if (n == 0) return (REAL)0;
    res = a[0] * b[0];
    for (int64_t i = 1; i < n; ++i) res = FMA(a[i], b[i], res);
  }
  return res;
}

/* operator selected by ConjugateGradient.h:202-218 from the UpLo template flag */
static void NAME(apply_op)(int64_t n, const int32_t* rowptr, const int32_t* colidx, const REAL* vals, int uplo,
                           const REAL* x, REAL* y) {
  if (uplo == 3)
    NAME(oracle_spmv)(n, rowptr, colidx, vals, x, y);
  else
    NAME(oracle_symv)(n, rowptr, colidx, vals, x, y, uplo);
}

/* ---- preconditioner hook -----------------------------------------------------------------------------------------
 * z = precond.solve(r) of the solver loops.  Jacobi / identity: z = invdiag .* r (BasicPreconditioners.h:88-101);
 * incomplete factorizations: the staged triangular solves of oracle.c (oracle_factors_apply, double only). */
typedef void (*NAME(oracle_precond_fn))(void* ctx, int64_t n, const REAL* r, REAL* z);
static void NAME(jacobi_apply)(void* ctx, int64_t n, const REAL* r, REAL* z) {
  const REAL* invdiag = (const REAL*)ctx;
  for (int64_t i = 0; i < n; ++i) z[i] = invdiag[i] * r[i];
}

/* ---- preconditioned conjugate gradient --------------------------------------------------------------------------
 * IterativeLinearSolvers/ConjugateGradient.h:26-91 (loop) and :197-221 (operator view, iterations/error/info).
 * x holds the initial guess on entry (zeros for solve(), IterativeSolverBase.h:399-404).
 * max_iters < 0 -> default 2*n (IterativeSolverBase.h:281-284); tol < 0 -> epsilon (:413).
 * info: 0 Success, 2 NoConvergence (Core/util/Constants.h:430-440). */
void NAME(oracle_cg_precond)(int64_t n, const int32_t* rowptr, const int32_t* colidx, const REAL* vals, const REAL* b,
                             REAL* x, REAL tol, int64_t max_iters, int uplo, int lanes, NAME(oracle_precond_fn) precond,
                             void* ctx, int64_t* iters_out, REAL* error_out, int* info_out) {
  if (max_iters < 0) max_iters = 2 * n;
  if (tol < 0) tol = EPS;
  REAL* r = (REAL*)malloc(sizeof(REAL) * (size_t)(n ? n : 1));
  REAL* p = (REAL*)malloc(sizeof(REAL) * (size_t)(n ? n : 1));
  REAL* z = (REAL*)malloc(sizeof(REAL) * (size_t)(n ? n : 1));
  REAL* tmp = (REAL*)malloc(sizeof(REAL) * (size_t)(n ? n : 1));

  int64_t it = 0;
  REAL err;
  NAME(apply_op)(n, rowptr, colidx, vals, uplo, x, tmp); /* :43 residual = rhs - mat*x */
  for (int64_t i = 0; i < n; ++i) r[i] = b[i] - tmp[i];
  REAL bb = NAME(oracle_dot)(b, b, n, lanes); /* :45 */
  if (bb == (REAL)0) {                        /* :46-52 */
    for (int64_t i = 0; i < n; ++i) x[i] = (REAL)0;
    it = 0;
    err = (REAL)0;
    goto done;
  }
  {
    REAL thr = tol * tol * bb; /* :53-54 */
    if (thr < TINY) thr = TINY;
    REAL rr = NAME(oracle_dot)(r, r, n, lanes); /* :55 */
    if (rr < thr) {                             /* :56-61 */
      it = 0;
      err = SQRT(rr / bb);
      goto done;
    }
    precond(ctx, n, r, p);                                    /* :63-64 */
    REAL abs_new = NAME(oracle_dot)(r, p, n, lanes);          /* :67 */
    while (it < max_iters) {                                  /* :69 */
      NAME(apply_op)(n, rowptr, colidx, vals, uplo, p, tmp);  /* :71 */
      REAL alpha = abs_new / NAME(oracle_dot)(p, tmp, n, lanes); /* :73 */
      for (int64_t i = 0; i < n; ++i) x[i] = FMA(alpha, p[i], x[i]);    /* :74 */
      for (int64_t i = 0; i < n; ++i) r[i] = FMA(-alpha, tmp[i], r[i]); /* :75 */
      rr = NAME(oracle_dot)(r, r, n, lanes);                            /* :77 */
      if (rr < thr) break;                                              /* :78-79, `it` not incremented */
      precond(ctx, n, r, z);                                            /* :81 */
      REAL abs_old = abs_new;
      abs_new = NAME(oracle_dot)(r, z, n, lanes); /* :84 */
      REAL beta = abs_new / abs_old;              /* :85 */
      for (int64_t i = 0; i < n; ++i) p[i] = FMA(beta, p[i], z[i]); /* :86 */
      ++it;
    }
    err = SQRT(rr / bb); /* :89 */
  }
done:
  if (iters_out) *iters_out = it;
  if (error_out) *error_out = err;
  if (info_out) *info_out = (err <= tol) ? 0 : 2; /* :220 */
  free(r); free(p); free(z); free(tmp);
}

/* Jacobi (precond 1) / identity (0): DiagonalPreconditioner / IdentityPreconditioner, BasicPreconditioners.h:64-101, :200-222 */
void NAME(oracle_cg)(int64_t n, const int32_t* rowptr, const int32_t* colidx, const REAL* vals, const REAL* b,
                     REAL* x, REAL tol, int64_t max_iters, int uplo, int precond, int lanes, int64_t* iters_out,
                     REAL* error_out, int* info_out) {
  REAL* invdiag = (REAL*)malloc(sizeof(REAL) * (size_t)(n ? n : 1));
  if (precond == 1)
    NAME(oracle_jacobi_factorize)(n, rowptr, colidx, vals, invdiag);
  else
    for (int64_t i = 0; i < n; ++i) invdiag[i] = (REAL)1;
  NAME(oracle_cg_precond)(n, rowptr, colidx, vals, b, x, tol, max_iters, uplo, lanes, NAME(jacobi_apply), invdiag,
                          iters_out, error_out, info_out);
  free(invdiag);
}

/* ---- preconditioned BiCGSTAB ---------------------------------------------------------------------------------------
 * IterativeLinearSolvers/BiCGSTAB.h:28-107 (loop) and :193-204 (info).  The operator is always the matrix as
 * stored.  When ||b|| == 0 the reference returns before touching iters/tol_error, so iterations() stays
 * maxIterations() and error() stays the tolerance (:47-51 with :196-199). */
void NAME(oracle_bicgstab_precond)(int64_t n, const int32_t* rowptr, const int32_t* colidx, const REAL* vals,
                                   const REAL* b, REAL* x, REAL tol, int64_t max_iters, int lanes,
                                   NAME(oracle_precond_fn) precond, void* ctx, int64_t* iters_out, REAL* error_out,
                                   int* info_out) {
  if (max_iters < 0) max_iters = 2 * n;
  if (tol < 0) tol = EPS;
  size_t bytes = sizeof(REAL) * (size_t)(n ? n : 1);
  REAL *r = (REAL*)malloc(bytes), *r0 = (REAL*)malloc(bytes), *v = (REAL*)calloc(n ? n : 1, sizeof(REAL));
  REAL *p = (REAL*)calloc(n ? n : 1, sizeof(REAL)), *y = (REAL*)malloc(bytes), *z = (REAL*)malloc(bytes);
  REAL *s = (REAL*)malloc(bytes), *t = (REAL*)malloc(bytes);

  int64_t it = max_iters;
  REAL err = tol;
  oracle_last_restarts = 0;
  NAME(oracle_spmv)(n, rowptr, colidx, vals, x, t); /* :42 r = rhs - mat*x */
  for (int64_t i = 0; i < n; ++i) r[i] = b[i] - t[i];
  memcpy(r0, r, sizeof(REAL) * (size_t)n);
  REAL r0_sqnorm = NAME(oracle_dot)(r0, r0, n, lanes);
  REAL rhs_sqnorm = NAME(oracle_dot)(b, b, n, lanes);
  if (rhs_sqnorm == (REAL)0) { /* :47-51 */
    for (int64_t i = 0; i < n; ++i) x[i] = (REAL)0;
    goto done;
  }
  {
    REAL rho = 1, alpha = 1, w = 1;
    REAL tol2 = tol * tol * rhs_sqnorm; /* :62 */
    REAL eps2 = EPS * EPS;              /* :63 */
    int64_t i_it = 0, restarts = 0;
    REAL rr = NAME(oracle_dot)(r, r, n, lanes);
    while (rr > tol2 && i_it < max_iters) { /* :67 */
      REAL rho_old = rho;
      rho = NAME(oracle_dot)(r0, r, n, lanes);    /* :71 */
      if (FABS(rho) < eps2 * r0_sqnorm) {         /* :72-81 restart */
        NAME(oracle_spmv)(n, rowptr, colidx, vals, x, t);
        for (int64_t i = 0; i < n; ++i) r[i] = b[i] - t[i];
        memcpy(r0, r, sizeof(REAL) * (size_t)n);
        rho = r0_sqnorm = NAME(oracle_dot)(r, r, n, lanes);
        if (restarts++ == 0) i_it = 0;
        oracle_last_restarts = restarts;
      }
      REAL beta = (rho / rho_old) * (alpha / w); /* :82 */
      for (int64_t i = 0; i < n; ++i) p[i] = FMA(beta, FMA(-w, v[i], p[i]), r[i]); /* :83 */
      precond(ctx, n, p, y);                                                       /* :85 */
      NAME(oracle_spmv)(n, rowptr, colidx, vals, y, v);                            /* :87 */
      alpha = rho / NAME(oracle_dot)(r0, v, n, lanes);                             /* :89 */
      for (int64_t i = 0; i < n; ++i) s[i] = FMA(-alpha, v[i], r[i]);              /* :90 */
      precond(ctx, n, s, z);                                                       /* :92 */
      NAME(oracle_spmv)(n, rowptr, colidx, vals, z, t);                            /* :93 */
      REAL tt = NAME(oracle_dot)(t, t, n, lanes);                                  /* :95 */
      if (tt > (REAL)0)
        w = NAME(oracle_dot)(t, s, n, lanes) / tt; /* :96-97 */
      else
        w = (REAL)0;
      /* :100 x += alpha*y + w*z: GCC fuses w*z into the inner add.  (Pinned bit-for-bit on the AVX-512 build of
       * oracle/_ref; the AVX2 build contracts the scalar remainder of this one statement differently, so on vectors
       * whose length is not a multiple of 4 its x differs from this port in the last bit.) */
      for (int64_t i = 0; i < n; ++i) x[i] = x[i] + FMA(w, z[i], alpha * y[i]);
      for (int64_t i = 0; i < n; ++i) r[i] = FMA(-w, t[i], s[i]);               /* :101 */
      ++i_it;
      rr = NAME(oracle_dot)(r, r, n, lanes);
    }
    err = SQRT(rr / rhs_sqnorm); /* :104 */
    it = i_it;
  }
done:
  if (iters_out) *iters_out = it;
  if (error_out) *error_out = err;
  if (info_out) *info_out = (err <= tol) ? 0 : 2; /* :201-203 */
  free(r); free(r0); free(v); free(p); free(y); free(z); free(s); free(t);
}

void NAME(oracle_bicgstab)(int64_t n, const int32_t* rowptr, const int32_t* colidx, const REAL* vals, const REAL* b,
                           REAL* x, REAL tol, int64_t max_iters, int precond, int lanes, int64_t* iters_out,
                           REAL* error_out, int* info_out) {
  REAL* invdiag = (REAL*)malloc(sizeof(REAL) * (size_t)(n ? n : 1));
  if (precond == 1)
    NAME(oracle_jacobi_factorize)(n, rowptr, colidx, vals, invdiag);
  else
    for (int64_t i = 0; i < n; ++i) invdiag[i] = (REAL)1;
  NAME(oracle_bicgstab_precond)(n, rowptr, colidx, vals, b, x, tol, max_iters, lanes, NAME(jacobi_apply), invdiag,
                                iters_out, error_out, info_out);
  free(invdiag);
}

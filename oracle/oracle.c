/* oracle/oracle.c -- TEST INFRASTRUCTURE: the CPU oracle ("port") for the sparse iterative-solve hot path.
 *
 * Plain-C restatement of Eigen's CSR SpMV, DiagonalPreconditioner, ConjugateGradient and BiCGSTAB loops
 * (see oracle_body.h for the per-function reference citations).  PARITY IS PINNED: with `lanes` matching the
 * ISA of oracle/_ref (the unmodified reference compiled from /root/reference by oracle/Makefile) every function
 * here is bit-identical to the reference on the committed golden vectors (tests/golden, tests/test_oracle_*.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library; the product
 * (eigen-git-mirror_b200/) never links, imports or calls it and has no CPU fallback.
 * Build: make -C oracle port  ->  oracle/_build/liboracle.so   (all FMAs are explicit; -ffp-contract=off)
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* How often the last oracle_bicgstab_* call took the re-orthogonalisation branch (BiCGSTAB.h:72-81); the reference
 * keeps this count in a local variable, tests need it to know that a case really exercises the branch. */
int64_t oracle_last_restarts = 0;

#define REAL double
#define FMA fma
#define SQRT sqrt
#define FABS fabs
#define EPS DBL_EPSILON
#define TINY DBL_MIN
#define NAME(x) x##_f64
#include "oracle_body.h"
#undef REAL
#undef FMA
#undef SQRT
#undef FABS
#undef EPS
#undef TINY
#undef NAME

#define REAL float
#define FMA fmaf
#define SQRT sqrtf
#define FABS fabsf
#define EPS FLT_EPSILON
#define TINY FLT_MIN
#define NAME(x) x##_f32
#include "oracle_body.h"

/* elementwise helpers used by tests: true residual ||b - A x|| / ||b|| in plain left-to-right double */
double oracle_true_residual_f64(int64_t n, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                                const double* x, const double* b) {
  double num = 0, den = 0;
  for (int64_t i = 0; i < n; ++i) {
    double t = 0;
    for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) t = fma(vals[k], x[colidx[k]], t);
    double d = b[i] - t;
    num += d * d;
    den += b[i] * b[i];
  }
  return den > 0 ? sqrt(num / den) : sqrt(num);
}

/* ---- incomplete-factorization preconditioners ----------------------------------------------
 * The reference applies IncompleteLUT / IncompleteCholesky with sequential substitutions (SparseCore/
 * TriangularSolver.h:26-134).  oracle_tri_stage restates ONE such substitution in the row-wise form the device uses
 * (eigen-git-mirror_b200/csrc/kernels_tri.cuh): row i: t = x[i]; t -= vals[k] * x[colidx[k]] over the row's entries in
 * storage order; x[i] = diag ? t / diag[i] : t.  Rows are visited in the order `order` lists them (n entries): the
 * natural order (ascending for a lower, descending for an upper factor) IS the reference's loop; the device's level
 * order must give the same bits because a row only reads rows of earlier levels.  fused != 0: each step is one FMA
 * -- what g++ -O3 emits for the column sweep `other.coeffRef(it.index(),col) -= tmp * it.value()` (:129) with FMA
 * available; fused == 0: product and subtraction rounded separately -- what it emits for the row-wise loops
 * `tmp -= lastVal * other.coeff(lastIndex,col)` (:49, :91).  Both probed against oracle/_ref (tests/test_factors.py). */
void oracle_tri_stage_f64(int64_t n, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                          const double* diag, const int32_t* order, int fused, double* x) {
  for (int64_t s = 0; s < n; ++s) {
    const int32_t i = order[s];
    double t = x[i];
    for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
      if (fused) t = fma(-vals[k], x[colidx[k]], t);
      else t = t - vals[k] * x[colidx[k]];
    }
    x[i] = diag ? t / diag[i] : t;
  }
}

/* out[k] = scale[k] * in[gather[k]]  (NULL gather = identity, NULL scale = 1): the permutation / scaling steps of
 * IncompleteLUT.h:172,175 and IncompleteCholesky.h:152-156 as the device performs them. */
void oracle_permute_scale_f64(int64_t n, const double* in, const int32_t* gather, const double* scale, double* out) {
  for (int64_t k = 0; k < n; ++k) {
    const double v = in[gather ? gather[k] : k];
    out[k] = scale ? scale[k] * v : v;
  }
}

/* A preconditioner object for oracle_cg_precond_f64 / oracle_bicgstab_precond_f64: the staged form of an incomplete
 * factorization (what eigen-git-mirror_b200/csrc/factors.cpp hands to the device), applied with the reference's natural
 * row order -- i.e. IncompleteLUT::_solve_impl (IncompleteLUT.h:171-176) / IncompleteCholesky::_solve_impl
 * (IncompleteCholesky.h:149-157) as restated by oracle_permute_scale_f64 + oracle_tri_stage_f64 above. */
typedef struct {
  int64_t n;
  const int32_t* pre_gather;   /* NULL = identity */
  const double* pre_scale;     /* NULL = 1 */
  const int32_t* post_gather;
  const double* post_scale;
  const int32_t* rowptr[2];
  const int32_t* colidx[2];
  const double* vals[2];
  const double* diag[2];       /* NULL = unit diagonal */
  const int32_t* order[2];     /* rows in the order they are solved */
  int32_t fused[2];
  double* work;                /* n doubles */
} oracle_factors;

void oracle_factors_apply(void* ctx, int64_t n, const double* r, double* z) {
  const oracle_factors* f = (const oracle_factors*)ctx;
  oracle_permute_scale_f64(n, r, f->pre_gather, f->pre_scale, f->work);
  for (int s = 0; s < 2; ++s)
    oracle_tri_stage_f64(n, f->rowptr[s], f->colidx[s], f->vals[s], f->diag[s], f->order[s], f->fused[s], f->work);
  oracle_permute_scale_f64(n, f->work, f->post_gather, f->post_scale, z);
}

/* ConjugateGradient<_, UpLo, IncompleteCholesky<...>> / BiCGSTAB<_, IncompleteLUT> with the loops of oracle_body.h */
void oracle_cg_factors_f64(int64_t n, const int32_t* rowptr, const int32_t* colidx, const double* vals, const double* b,
                           double* x, double tol, int64_t max_iters, int uplo, int lanes, oracle_factors* f,
                           int64_t* iters_out, double* error_out, int* info_out) {
  oracle_cg_precond_f64(n, rowptr, colidx, vals, b, x, tol, max_iters, uplo, lanes, oracle_factors_apply, f, iters_out,
                        error_out, info_out);
}
void oracle_bicgstab_factors_f64(int64_t n, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                                 const double* b, double* x, double tol, int64_t max_iters, int lanes, oracle_factors* f,
                                 int64_t* iters_out, double* error_out, int* info_out) {
  oracle_bicgstab_precond_f64(n, rowptr, colidx, vals, b, x, tol, max_iters, lanes, oracle_factors_apply, f, iters_out,
                              error_out, info_out);
}

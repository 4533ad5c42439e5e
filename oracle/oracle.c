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

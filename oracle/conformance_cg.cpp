// oracle/conformance_cg.cpp -- TEST INFRASTRUCTURE (API conformance, runs on the GPU box).
//
// The reference's OWN solver test driver (test/sparse_solver.h: check_sparse_spd_solving, :273-356, which runs
// check_sparse_solving :41-145 over dense / sparse / multi-column right-hand sides, solveWithGuess,
// analyzePattern+factorize, Map, uncompressed and expression inputs) instantiated on b200::ConjugateGradient, with
// the instantiation list of test/conjugate_gradient.cpp:13-34 restricted to what the B200 path supports (double;
// Jacobi and identity preconditioners; int and long storage indices).  Compiled in the build container against the
// reference headers where they lie (oracle/Makefile target `conformance`), linked to libb200sparse.so, and executed
// on the GPU box by tests/test_gpu_conformance.py.
#include "sparse_solver.h"

#include <b200/IterativeSolvers.h>

template <typename T, typename I_>
void test_b200_conjugate_gradient_T() {
  typedef SparseMatrix<T, 0, I_> SparseMatrixType;
  b200::ConjugateGradient<SparseMatrixType, Lower> cg_colmajor_lower_diag;
  b200::ConjugateGradient<SparseMatrixType, Upper> cg_colmajor_upper_diag;
  b200::ConjugateGradient<SparseMatrixType, Lower | Upper> cg_colmajor_loup_diag;
  b200::ConjugateGradient<SparseMatrixType, Lower, IdentityPreconditioner> cg_colmajor_lower_I;
  b200::ConjugateGradient<SparseMatrixType, Upper, IdentityPreconditioner> cg_colmajor_upper_I;

  CALL_SUBTEST(check_sparse_spd_solving(cg_colmajor_lower_diag));
  CALL_SUBTEST(check_sparse_spd_solving(cg_colmajor_upper_diag));
  CALL_SUBTEST(check_sparse_spd_solving(cg_colmajor_loup_diag));
  CALL_SUBTEST(check_sparse_spd_solving(cg_colmajor_lower_I));
  CALL_SUBTEST(check_sparse_spd_solving(cg_colmajor_upper_I));

  typedef SparseMatrix<T, RowMajor, I_> RowMajorType;
  b200::ConjugateGradient<RowMajorType, Lower | Upper> cg_rowmajor_loup_diag;
  b200::ConjugateGradient<RowMajorType, Lower> cg_rowmajor_lower_diag;
  CALL_SUBTEST(check_sparse_spd_solving(cg_rowmajor_loup_diag));
  CALL_SUBTEST(check_sparse_spd_solving(cg_rowmajor_lower_diag));
}

EIGEN_DECLARE_TEST(b200_conjugate_gradient) {
  CALL_SUBTEST_1((test_b200_conjugate_gradient_T<double, int>()));
  CALL_SUBTEST_1((test_b200_conjugate_gradient_T<double, long int>()));
}

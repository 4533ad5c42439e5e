// oracle/conformance_bicgstab.cpp -- TEST INFRASTRUCTURE (API conformance, runs on the GPU box).
//
// The reference's own driver check_sparse_square_solving (test/sparse_solver.h:403-478) instantiated on
// b200::BiCGSTAB, following test/bicgstab.cpp:13-34 for the preconditioners the B200 path supports (Jacobi, identity;
// ILUT is out of scope).  See conformance_cg.cpp for how it is built and run.
#include "sparse_solver.h"

#include <b200/IterativeSolvers.h>

template <typename T, typename I_>
void test_b200_bicgstab_T() {
  b200::BiCGSTAB<SparseMatrix<T, 0, I_>, DiagonalPreconditioner<T> > bicgstab_colmajor_diag;
  b200::BiCGSTAB<SparseMatrix<T, 0, I_>, IdentityPreconditioner> bicgstab_colmajor_I;
  b200::BiCGSTAB<SparseMatrix<T, RowMajor, I_>, DiagonalPreconditioner<T> > bicgstab_rowmajor_diag;

  bicgstab_colmajor_diag.setTolerance(NumTraits<T>::epsilon() * 4);
  bicgstab_colmajor_I.setTolerance(NumTraits<T>::epsilon() * 4);
  bicgstab_rowmajor_diag.setTolerance(NumTraits<T>::epsilon() * 4);

  CALL_SUBTEST(check_sparse_square_solving(bicgstab_colmajor_diag));
  CALL_SUBTEST(check_sparse_square_solving(bicgstab_colmajor_I));
  CALL_SUBTEST(check_sparse_square_solving(bicgstab_rowmajor_diag));
}

EIGEN_DECLARE_TEST(b200_bicgstab) {
  CALL_SUBTEST_1((test_b200_bicgstab_T<double, int>()));
  CALL_SUBTEST_1((test_b200_bicgstab_T<double, long int>()));
}

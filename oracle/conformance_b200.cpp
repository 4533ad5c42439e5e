// oracle/conformance_b200.cpp -- TEST INFRASTRUCTURE (API conformance; built here, executed on the GPU box).
//
// Instantiates the reference's OWN solver test drivers on the B200 binding classes:
//   check_sparse_spd_solving     test/sparse_solver.h:273-356  -> b200::ConjugateGradient
//   check_sparse_square_solving  test/sparse_solver.h:403-478  -> b200::BiCGSTAB
// Both run check_sparse_solving (:41-145) over dense, sparse and multi-column right-hand sides, solveWithGuess,
// analyzePattern + factorize, Map / uncompressed / expression inputs and the matrix constructor, and compare with a
// dense Householder-QR solve at the reference's own tolerance.  The solver types cover what the reference's
// conjugate_gradient / bicgstab tests cover minus what the B200 path does not implement (complex scalars, ILUT).
// Built by `make -C oracle conformance` against the reference headers where they lie; linked to libb200sparse.so.
#include "sparse_solver.h"

#include <b200/IterativeSolvers.h>

namespace {

template <typename Index_> using ColMat = SparseMatrix<double, ColMajor, Index_>;
template <typename Index_> using RowMat = SparseMatrix<double, RowMajor, Index_>;

template <typename Solver>
void spd_case() {
  Solver solver;
  CALL_SUBTEST(check_sparse_spd_solving(solver));
}

template <typename Solver>
void square_case() {
  Solver solver;
  solver.setTolerance(4 * NumTraits<double>::epsilon());
  CALL_SUBTEST(check_sparse_square_solving(solver));
}

template <typename Index_>
void cg_suite() {
  typedef DiagonalPreconditioner<double> Jacobi;
  spd_case<b200::ConjugateGradient<ColMat<Index_>, Lower, Jacobi> >();
  spd_case<b200::ConjugateGradient<ColMat<Index_>, Upper, Jacobi> >();
  spd_case<b200::ConjugateGradient<ColMat<Index_>, Lower | Upper, Jacobi> >();
  spd_case<b200::ConjugateGradient<ColMat<Index_>, Lower, IdentityPreconditioner> >();
  spd_case<b200::ConjugateGradient<ColMat<Index_>, Upper, IdentityPreconditioner> >();
  spd_case<b200::ConjugateGradient<RowMat<Index_>, Lower | Upper, Jacobi> >();
  spd_case<b200::ConjugateGradient<RowMat<Index_>, Lower, Jacobi> >();
}

template <typename Index_>
void bicgstab_suite() {
  square_case<b200::BiCGSTAB<ColMat<Index_>, DiagonalPreconditioner<double> > >();
  square_case<b200::BiCGSTAB<ColMat<Index_>, IdentityPreconditioner> >();
  square_case<b200::BiCGSTAB<RowMat<Index_>, DiagonalPreconditioner<double> > >();
}

}  // namespace

EIGEN_DECLARE_TEST(b200_conjugate_gradient) {
  CALL_SUBTEST_1(cg_suite<int>());
  CALL_SUBTEST_1(cg_suite<long int>());
}

EIGEN_DECLARE_TEST(b200_bicgstab) {
  CALL_SUBTEST_1(bicgstab_suite<int>());
  CALL_SUBTEST_1(bicgstab_suite<long int>());
}

// oracle/conformance_precond_b200.cpp -- TEST INFRASTRUCTURE (API conformance; built here, executed on the GPU box).
//
// The reference's own test drivers for solvers with incomplete-factorization preconditioners, instantiated on the B200
// binding classes:
//   test/incomplete_cholesky.cpp:15-31   ConjugateGradient<_, UpLo, IncompleteCholesky<T, UpLo, AMD | Natural>>  (+ bug 1150, :34-61)
//   test/bicgstab.cpp:17,25              BiCGSTAB<_, IncompleteLUT<T, I>>
//   unsupported/test/gmres.cpp:18,23     GMRES<_, IncompleteLUT<T>>
// check_sparse_spd_solving / check_sparse_square_solving (test/sparse_solver.h) run dense, sparse and multi-column
// right-hand sides, solveWithGuess, analyzePattern + factorize, Map / uncompressed / expression inputs, and compare with
// a dense solve at the reference's tolerance.  The preconditioner object factorizes on the host with the reference's own
// code (it is the solver's m_preconditioner, IterativeSolverBase.h:196-247); its application in every iteration runs on
// the GPU.  Kept apart from conformance_b200.cpp so that the established drivers keep their own binary.
// Built by `make -C oracle conformance` against the reference headers where they lie; linked to libb200sparse.so.
#include "sparse_solver.h"

#include <b200/IterativeSolvers.h>
#include <b200/KrylovSolvers.h>
#include <b200/Ordering.h>

namespace {

template <typename T, typename I_>
void incomplete_cholesky_suite() {
  typedef SparseMatrix<T, 0, I_> SparseMatrixType;
  b200::ConjugateGradient<SparseMatrixType, Lower, IncompleteCholesky<T, Lower, AMDOrdering<I_> > > cg_illt_lower_amd;
  b200::ConjugateGradient<SparseMatrixType, Lower, IncompleteCholesky<T, Lower, NaturalOrdering<I_> > > cg_illt_lower_nat;
  b200::ConjugateGradient<SparseMatrixType, Upper, IncompleteCholesky<T, Upper, AMDOrdering<I_> > > cg_illt_upper_amd;
  b200::ConjugateGradient<SparseMatrixType, Upper, IncompleteCholesky<T, Upper, NaturalOrdering<I_> > > cg_illt_upper_nat;
  b200::ConjugateGradient<SparseMatrixType, Upper | Lower, IncompleteCholesky<T, Lower, AMDOrdering<I_> > > cg_illt_uplo_amd;
  CALL_SUBTEST(check_sparse_spd_solving(cg_illt_lower_amd));
  CALL_SUBTEST(check_sparse_spd_solving(cg_illt_lower_nat));
  CALL_SUBTEST(check_sparse_spd_solving(cg_illt_upper_amd));
  CALL_SUBTEST(check_sparse_spd_solving(cg_illt_upper_nat));
  CALL_SUBTEST(check_sparse_spd_solving(cg_illt_uplo_amd));
}

// the ordering made for the GPU (include/b200/Ordering.h) in the same slot as AMDOrdering / NaturalOrdering
template <typename T, typename I_>
void multicolor_suite() {
  typedef SparseMatrix<T, 0, I_> SparseMatrixType;
  b200::ConjugateGradient<SparseMatrixType, Lower, IncompleteCholesky<T, Lower, b200::MulticolorOrdering<I_> > > cg_illt_lower_mc;
  b200::ConjugateGradient<SparseMatrixType, Upper, IncompleteCholesky<T, Upper, b200::MulticolorOrdering<I_> > > cg_illt_upper_mc;
  CALL_SUBTEST(check_sparse_spd_solving(cg_illt_lower_mc));
  CALL_SUBTEST(check_sparse_spd_solving(cg_illt_upper_mc));
}

void bug1150() {  // test/incomplete_cholesky.cpp:34-61
  for (int N = 1; N < 20; ++N) {
    Eigen::MatrixXd b(N, N);
    b.setOnes();
    Eigen::SparseMatrix<double> m(N, N);
    m.reserve(Eigen::VectorXi::Constant(N, 4));
    for (int i = 0; i < N; ++i) {
      m.insert(i, i) = 1;
      m.coeffRef(i, i / 2) = 2;
      m.coeffRef(i, i / 3) = 2;
      m.coeffRef(i, i / 4) = 2;
    }
    Eigen::SparseMatrix<double> A;
    A = m * m.transpose();
    b200::ConjugateGradient<Eigen::SparseMatrix<double>, Eigen::Lower | Eigen::Upper, Eigen::IncompleteCholesky<double> > solver(A);
    VERIFY(solver.preconditioner().info() == Eigen::Success);
    VERIFY(solver.info() == Eigen::Success);
  }
}

template <typename T, typename I_>
void ilut_suite() {
  b200::BiCGSTAB<SparseMatrix<T, 0, I_>, IncompleteLUT<T, I_> > bicgstab_colmajor_ilut;
  bicgstab_colmajor_ilut.setTolerance(NumTraits<T>::epsilon() * 4);  // test/bicgstab.cpp:21
  CALL_SUBTEST(check_sparse_square_solving(bicgstab_colmajor_ilut));
  b200::BiCGSTAB<SparseMatrix<T, RowMajor, I_>, IncompleteLUT<T, I_> > bicgstab_rowmajor_ilut;
  bicgstab_rowmajor_ilut.setTolerance(NumTraits<T>::epsilon() * 4);
  CALL_SUBTEST(check_sparse_square_solving(bicgstab_rowmajor_ilut));
}

}  // namespace

#ifdef B200_CONFORMANCE_ORDERING_ONLY  // built as its own binary (conformance_ordering_b200)
EIGEN_DECLARE_TEST(b200_multicolor_ordering) {
  CALL_SUBTEST_1((multicolor_suite<double, int>()));
  CALL_SUBTEST_1((multicolor_suite<double, long int>()));
}
#else
EIGEN_DECLARE_TEST(b200_incomplete_cholesky) {
  CALL_SUBTEST_1((incomplete_cholesky_suite<double, int>()));
  CALL_SUBTEST_1((incomplete_cholesky_suite<double, long int>()));
  CALL_SUBTEST_1(bug1150());
}

EIGEN_DECLARE_TEST(b200_ilut) {
  CALL_SUBTEST_1((ilut_suite<double, int>()));
  CALL_SUBTEST_1((ilut_suite<double, long int>()));
  b200::GMRES<SparseMatrix<double>, IncompleteLUT<double> > gmres_colmajor_ilut;
  CALL_SUBTEST_1(check_sparse_square_solving(gmres_colmajor_ilut));
}
#endif

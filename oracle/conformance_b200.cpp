// oracle/conformance_b200.cpp -- TEST INFRASTRUCTURE (API conformance; built here, executed on the GPU box).
//
// Instantiates the reference's OWN solver test drivers on the B200 binding classes:
//   check_sparse_spd_solving     test/sparse_solver.h:273-356  -> b200::ConjugateGradient
//   check_sparse_square_solving  test/sparse_solver.h:403-478  -> b200::BiCGSTAB
// Both run check_sparse_solving (:41-145) over dense, sparse and multi-column right-hand sides, solveWithGuess,
// analyzePattern + factorize, Map / uncompressed / expression inputs and the matrix constructor, and compare with a
// dense Householder-QR solve at the reference's own tolerance.  The solver types cover what the reference's
// conjugate_gradient / bicgstab tests cover minus what the B200 path does not implement (complex scalars, ILUT),
// plus the float instantiations.  The SpMV drop-in (b200::SparseOperator) is put through the sparse * dense forms of
// test/sparse_product.cpp:140-159 and through the reference's own CPU ConjugateGradient as a matrix-free operator
// (doc/examples/matrixfree_cg.cpp).
// Built by `make -C oracle conformance` against the reference headers where they lie; linked to libb200sparse.so.
#include "sparse_solver.h"

#include "sparse.h"

#include <b200/IterativeSolvers.h>
#include <b200/KrylovSolvers.h>
#include <b200/SparseOperator.h>

namespace {

template <typename T, typename Index_> using ColMat = SparseMatrix<T, ColMajor, Index_>;
template <typename T, typename Index_> using RowMat = SparseMatrix<T, RowMajor, Index_>;

template <typename Solver>
void spd_case() {
  Solver solver;
  CALL_SUBTEST(check_sparse_spd_solving(solver));
}

template <typename Solver>
void square_case() {
  Solver solver;
  solver.setTolerance(4 * NumTraits<typename Solver::Scalar>::epsilon());
  CALL_SUBTEST(check_sparse_square_solving(solver));
}

template <typename T, typename Index_>
void cg_suite() {
  typedef DiagonalPreconditioner<T> Jacobi;
  spd_case<b200::ConjugateGradient<ColMat<T, Index_>, Lower, Jacobi> >();
  spd_case<b200::ConjugateGradient<ColMat<T, Index_>, Upper, Jacobi> >();
  spd_case<b200::ConjugateGradient<ColMat<T, Index_>, Lower | Upper, Jacobi> >();
  spd_case<b200::ConjugateGradient<ColMat<T, Index_>, Lower, IdentityPreconditioner> >();
  spd_case<b200::ConjugateGradient<ColMat<T, Index_>, Upper, IdentityPreconditioner> >();
  spd_case<b200::ConjugateGradient<RowMat<T, Index_>, Lower | Upper, Jacobi> >();
  spd_case<b200::ConjugateGradient<RowMat<T, Index_>, Lower, Jacobi> >();
}

template <typename T, typename Index_>
void bicgstab_suite() {
  square_case<b200::BiCGSTAB<ColMat<T, Index_>, DiagonalPreconditioner<T> > >();
  square_case<b200::BiCGSTAB<ColMat<T, Index_>, IdentityPreconditioner> >();
  square_case<b200::BiCGSTAB<RowMat<T, Index_>, DiagonalPreconditioner<T> > >();
}

// The sparse * dense forms of test/sparse_product.cpp:140-159 with the device operator in place of the sparse matrix.
template <typename Scalar>
void sparse_operator_products() {
  typedef Matrix<Scalar, Dynamic, Dynamic> DenseMatrix;
  typedef Matrix<Scalar, Dynamic, 1> DenseVector;
  const Index rows = internal::random<Index>(1, 200), depth = internal::random<Index>(1, 200),
              cols = internal::random<Index>(1, 8);
  const double density = (std::max)(8. / (rows * depth), 0.2);
  DenseMatrix refMat2 = DenseMatrix::Zero(rows, depth);
  SparseMatrix<Scalar, RowMajor> m2(rows, depth);
  initSparse<Scalar>(density, refMat2, m2);
  SparseMatrix<Scalar, ColMajor> m2c = m2;
  DenseMatrix refMat3 = DenseMatrix::Random(depth, cols), refMat5 = DenseMatrix::Random(depth, cols);
  DenseMatrix refMat4 = DenseMatrix::Random(rows, cols), dm4 = refMat4;
  DenseMatrix refMat3t = refMat3.transpose();

  b200::SparseOperator<Scalar> op(m2);
  VERIFY(op.info() == Success);
  VERIFY(op.rows() == rows && op.cols() == depth);
  b200::SparseOperator<Scalar> opc(m2c);  // column-major input: converted once at compute()
  VERIFY(opc.info() == Success);

  // sparse * dense matrix
  VERIFY_IS_APPROX(dm4 = op * refMat3, refMat4 = refMat2 * refMat3);
  VERIFY_IS_APPROX(dm4 = opc * refMat3, refMat4 = refMat2 * refMat3);
  VERIFY_IS_APPROX(dm4 = op * refMat3t.transpose(), refMat4 = refMat2 * refMat3t.transpose());
  VERIFY_IS_APPROX(dm4 = dm4 + op * refMat3, refMat4 = refMat4 + refMat2 * refMat3);
  VERIFY_IS_APPROX(dm4 += op * refMat3, refMat4 += refMat2 * refMat3);
  VERIFY_IS_APPROX(dm4 -= op * refMat3, refMat4 -= refMat2 * refMat3);
  VERIFY_IS_APPROX(dm4.noalias() += op * refMat3, refMat4 += refMat2 * refMat3);
  VERIFY_IS_APPROX(dm4.noalias() -= op * refMat3, refMat4 -= refMat2 * refMat3);
  VERIFY_IS_APPROX(dm4 = op * (refMat3 + refMat3), refMat4 = refMat2 * (refMat3 + refMat3));
  VERIFY_IS_APPROX(dm4 = op * ((refMat3 + refMat5) * Scalar(0.5)), refMat4 = refMat2 * ((refMat3 + refMat5) * Scalar(0.5)));
  // sparse * dense vector
  VERIFY_IS_APPROX(dm4.col(0) = op * refMat3.col(0), refMat4.col(0) = refMat2 * refMat3.col(0));
  VERIFY_IS_APPROX(dm4.col(0) = op * refMat3t.transpose().col(0), refMat4.col(0) = refMat2 * refMat3t.transpose().col(0));
  DenseVector x = DenseVector::Random(depth), y(rows), yref(rows);
  y.noalias() = op * x;  // the exact statement of ConjugateGradient.h:71
  yref.noalias() = m2 * x;
  VERIFY_IS_APPROX(y, yref);
  // zero matrix (test/sparse_product.cpp:336-355) and 1x1 (bug_942, :357-378)
  SparseMatrix<Scalar, RowMajor> zero(rows, depth);
  b200::SparseOperator<Scalar> opz(zero);
  VERIFY(opz.info() == Success);
  VERIFY_IS_APPROX(dm4 = opz * refMat3, refMat4 = DenseMatrix::Zero(rows, cols));
  SparseMatrix<Scalar, RowMajor> one(1, 1);
  one.insert(0, 0) = Scalar(2);
  b200::SparseOperator<Scalar> op1(one);
  DenseVector d1 = DenseVector::Constant(1, Scalar(3)), r1 = op1 * d1;
  VERIFY_IS_APPROX(r1[0], Scalar(6));
}

// The device product as a matrix-free operator of the reference's own CPU solver (doc/examples/matrixfree_cg.cpp).
void sparse_operator_matrixfree_cg() {
  typedef SparseMatrix<double, RowMajor> SpMat;
  const int n = 40;
  SpMat A(n * n, n * n);
  std::vector<Triplet<double> > trip;
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      const int r = i + n * j;
      trip.push_back(Triplet<double>(r, r, 4.0));
      if (i > 0) trip.push_back(Triplet<double>(r, r - 1, -1.0));
      if (i < n - 1) trip.push_back(Triplet<double>(r, r + 1, -1.0));
      if (j > 0) trip.push_back(Triplet<double>(r, r - n, -1.0));
      if (j < n - 1) trip.push_back(Triplet<double>(r, r + n, -1.0));
    }
  A.setFromTriplets(trip.begin(), trip.end());
  VectorXd b = VectorXd::Random(n * n);
  b200::SparseOperator<double> op(A);
  VERIFY(op.info() == Success);
  Eigen::ConjugateGradient<b200::SparseOperator<double>, Lower | Upper, IdentityPreconditioner> mf;
  mf.compute(op);
  mf.setTolerance(1e-10);
  VectorXd x = mf.solve(b);
  Eigen::ConjugateGradient<SpMat, Lower | Upper, IdentityPreconditioner> ref(A);
  ref.setTolerance(1e-10);
  VectorXd xr = ref.solve(b);
  VERIFY(mf.info() == Success);
  VERIFY(mf.iterations() == ref.iterations());  // 5-point rows: one thread per row, bit-identical products
  VERIFY_IS_APPROX(x, xr);
}

}  // namespace

EIGEN_DECLARE_TEST(b200_conjugate_gradient) {
  CALL_SUBTEST_1((cg_suite<double, int>()));
  CALL_SUBTEST_1((cg_suite<double, long int>()));
  CALL_SUBTEST_1((cg_suite<float, int>()));
}

EIGEN_DECLARE_TEST(b200_bicgstab) {
  CALL_SUBTEST_1((bicgstab_suite<double, int>()));
  CALL_SUBTEST_1((bicgstab_suite<double, long int>()));
  CALL_SUBTEST_1((bicgstab_suite<float, int>()));
}

// test/lscg.cpp:13-32, unsupported/test/minres.cpp:16-38, unsupported/test/gmres.cpp:13-24 on the b200 classes
EIGEN_DECLARE_TEST(b200_krylov) {
  {
    b200::LeastSquaresConjugateGradient<SparseMatrix<double> > lscg_colmajor_diag;
    b200::LeastSquaresConjugateGradient<SparseMatrix<double>, IdentityPreconditioner> lscg_colmajor_I;
    b200::LeastSquaresConjugateGradient<SparseMatrix<double, RowMajor> > lscg_rowmajor_diag;
    b200::LeastSquaresConjugateGradient<SparseMatrix<double, RowMajor>, IdentityPreconditioner> lscg_rowmajor_I;
    CALL_SUBTEST_1(check_sparse_square_solving(lscg_colmajor_diag));
    CALL_SUBTEST_1(check_sparse_square_solving(lscg_colmajor_I));
    CALL_SUBTEST_1(check_sparse_leastsquare_solving(lscg_colmajor_diag));
    CALL_SUBTEST_1(check_sparse_leastsquare_solving(lscg_colmajor_I));
    CALL_SUBTEST_1(check_sparse_square_solving(lscg_rowmajor_diag));
    CALL_SUBTEST_1(check_sparse_square_solving(lscg_rowmajor_I));
    CALL_SUBTEST_1(check_sparse_leastsquare_solving(lscg_rowmajor_diag));
    CALL_SUBTEST_1(check_sparse_leastsquare_solving(lscg_rowmajor_I));
  }
  {
    b200::MINRES<SparseMatrix<double>, Lower, IdentityPreconditioner> minres_colmajor_lower_I;
    b200::MINRES<SparseMatrix<double>, Upper, IdentityPreconditioner> minres_colmajor_upper_I;
    b200::MINRES<SparseMatrix<double>, Lower, DiagonalPreconditioner<double> > minres_colmajor_lower_diag;
    b200::MINRES<SparseMatrix<double>, Upper, DiagonalPreconditioner<double> > minres_colmajor_upper_diag;
    b200::MINRES<SparseMatrix<double>, Lower | Upper, DiagonalPreconditioner<double> > minres_colmajor_uplo_diag;
    CALL_SUBTEST_1(check_sparse_spd_solving(minres_colmajor_lower_I));
    CALL_SUBTEST_1(check_sparse_spd_solving(minres_colmajor_upper_I));
    CALL_SUBTEST_1(check_sparse_spd_solving(minres_colmajor_lower_diag));
    CALL_SUBTEST_1(check_sparse_spd_solving(minres_colmajor_upper_diag));
    CALL_SUBTEST_1(check_sparse_spd_solving(minres_colmajor_uplo_diag));
  }
  {
    b200::GMRES<SparseMatrix<double>, DiagonalPreconditioner<double> > gmres_colmajor_diag;
    b200::GMRES<SparseMatrix<double, RowMajor>, DiagonalPreconditioner<double> > gmres_rowmajor_diag;
    CALL_SUBTEST_1(check_sparse_square_solving(gmres_colmajor_diag));
    CALL_SUBTEST_1(check_sparse_square_solving(gmres_rowmajor_diag));
  }
}

EIGEN_DECLARE_TEST(b200_sparse_operator) {
  for (int i = 0; i < g_repeat; i++) {
    CALL_SUBTEST_1(sparse_operator_products<double>());
    CALL_SUBTEST_1(sparse_operator_products<float>());
  }
  CALL_SUBTEST_1(sparse_operator_matrixfree_cg());
}

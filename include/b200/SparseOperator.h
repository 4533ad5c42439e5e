// b200/SparseOperator.h -- the SpMV-only drop-in: `y = A * x` of a row-major sparse matrix on a B200.
//
// Replaces, behind Eigen's own product expression machinery,
//   sparse_time_dense_product_impl<...,RowMajor,true>::run / processRow   Eigen/src/SparseCore/SparseDenseProduct.h:26-72
//   generic_product_impl<Lhs,Rhs,SparseShape,DenseShape>::scaleAndAddTo   Eigen/src/SparseCore/SparseDenseProduct.h:178-193
// through the hook the reference documents for user operators: a type that "looks like" a sparse matrix plus a
// specialisation of internal::generic_product_impl (doc/examples/matrixfree_cg.cpp:14-76, generic_matrix_wrapper
// IterativeSolverBase.h:98-133).  All forms the reference's own test exercises for sparse * dense
// (test/sparse_product.cpp:140-159: `=`, `+=`, `-=`, `.noalias() +=`, expression right-hand sides, vector and
// multi-column right-hand sides) go through Eigen's generic evaluators and end in scaleAndAddTo below, which calls
// b200s_spmv_f64 / _f32 (include/b200sparse.h) once per right-hand-side column.
//
// The operator also plugs into the reference's CPU solvers as a matrix-free operator, e.g.
//   Eigen::ConjugateGradient<b200::SparseOperator<double>, Eigen::Lower|Eigen::Upper, Eigen::IdentityPreconditioner>
// but that path moves x and y over PCIe every iteration: it is the SpMV drop-in and a parity tool, not the fast path
// (b200::ConjugateGradient keeps the whole iteration on the device).
//
// Row-partitioned use: setDistributed(...) before compute(); the matrix is this rank's rows_local x N block, x is this
// rank's block of the vector (rows_local entries) and so is y -- the halo is exchanged on the device.
#ifndef B200_SPARSE_OPERATOR_H
#define B200_SPARSE_OPERATOR_H

#include "IterativeSolvers.h"

namespace b200 {
template <typename Scalar_>
class SparseOperator;
}

namespace Eigen {
namespace internal {
// looks like a row-major sparse matrix (matrixfree_cg.cpp:11-17)
template <typename Scalar_>
struct traits<b200::SparseOperator<Scalar_> > : public traits<SparseMatrix<Scalar_, RowMajor, int> > {};
}  // namespace internal
}  // namespace Eigen

namespace b200 {

template <typename Scalar_>
class SparseOperator : public Eigen::EigenBase<SparseOperator<Scalar_> > {
 public:
  typedef Scalar_ Scalar;
  typedef typename Eigen::NumTraits<Scalar>::Real RealScalar;
  typedef int StorageIndex;
  typedef Eigen::Index Index;
  enum { ColsAtCompileTime = Eigen::Dynamic, MaxColsAtCompileTime = Eigen::Dynamic, IsRowMajor = true };

  EIGEN_STATIC_ASSERT(detail::abi<Scalar>::supported, THIS_TYPE_IS_NOT_SUPPORTED)

  SparseOperator() : m_rows(0), m_cols(0), m_ok(false) {}
  template <typename Derived>
  explicit SparseOperator(const Eigen::SparseMatrixBase<Derived>& A) : m_rows(0), m_cols(0), m_ok(false) {
    compute(A.derived());
  }

  SparseOperator& setDistributed(int rank, int world, const int64_t* row_starts, b200s_allgather_fn allgather,
                                 void* allgather_ctx, int device = -1) {
    b200s_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.device = device;
    cfg.rank = rank;
    cfg.world = world;
    cfg.allgather = allgather;
    cfg.allgather_ctx = allgather_ctx;
    m_dev.configure(cfg, row_starts);
    return *this;
  }

  // Uploads A (any storage order / expression: evaluated to compressed row-major once, as Ref<const SparseMatrix<...,
  // RowMajor>, StandardCompressedFormat> would).  A may be destroyed afterwards: the device holds the copy.
  template <typename Derived>
  SparseOperator& compute(const Eigen::SparseMatrixBase<Derived>& A) {
    Eigen::SparseMatrix<Scalar, Eigen::RowMajor, int> csr = A.derived();
    csr.makeCompressed();
    m_rows = csr.rows();
    m_cols = csr.cols();
    m_ok = m_dev.analyze(csr, B200S_BOTH, false) && m_dev.factorize(csr, B200S_PRECOND_IDENTITY);
    return *this;
  }

  Index rows() const { return m_rows; }
  Index cols() const { return m_cols; }
  Eigen::ComputationInfo info() const { return m_ok ? Eigen::Success : Eigen::InvalidInput; }
  const std::string& lastError() const { return m_dev.lastError(); }

  template <typename Rhs>
  Eigen::Product<SparseOperator, Rhs, Eigen::AliasFreeProduct> operator*(const Eigen::MatrixBase<Rhs>& x) const {
    return Eigen::Product<SparseOperator, Rhs, Eigen::AliasFreeProduct>(*this, x.derived());
  }

  // y = A x on contiguous host vectors (what scaleAndAddTo calls per column)
  bool multiply(const Scalar* x, Scalar* y) const { return m_ok && m_dev.multiply(x, y); }

 private:
  Index m_rows, m_cols;
  bool m_ok;
  mutable detail::DeviceSolver<Scalar> m_dev;
};

}  // namespace b200

namespace Eigen {
namespace internal {

// dst += alpha * op * rhs, one device product per right-hand-side column (SparseDenseProduct.h:185-192 is the
// reference's counterpart; evalTo = setZero + scaleAndAddTo comes from generic_product_impl_base, ProductEvaluators.h:348-349)
template <typename Scalar_, typename Rhs, int ProductTag>
struct generic_product_impl<b200::SparseOperator<Scalar_>, Rhs, SparseShape, DenseShape, ProductTag>
    : generic_product_impl_base<b200::SparseOperator<Scalar_>, Rhs,
                                generic_product_impl<b200::SparseOperator<Scalar_>, Rhs, SparseShape, DenseShape, ProductTag> > {
  typedef Scalar_ Scalar;

  template <typename Dest>
  static void scaleAndAddTo(Dest& dst, const b200::SparseOperator<Scalar_>& lhs, const Rhs& rhs, const Scalar& alpha) {
    typedef Matrix<Scalar, Dynamic, 1> Vec;
    // the right-hand side may be an expression, a strided block or a row-major matrix: evaluate column by column
    typename nested_eval<Rhs, Dynamic>::type actual_rhs(rhs);
    Vec x(actual_rhs.rows()), y(lhs.rows());
    for (Index c = 0; c < actual_rhs.cols(); ++c) {
      x = actual_rhs.col(c);
      if (!lhs.multiply(x.data(), y.data())) {
        // no CPU fallback: poison the result so that a failed product cannot pass for a computed one
        dst.col(c).setConstant(std::numeric_limits<Scalar>::quiet_NaN());
        continue;
      }
      dst.col(c) += alpha * y;
    }
  }
};

}  // namespace internal
}  // namespace Eigen

#endif  // B200_SPARSE_OPERATOR_H

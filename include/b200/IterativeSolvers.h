// b200/IterativeSolvers.h -- the reference-side binding: Eigen solver classes whose iteration runs on B200 GPUs.
//
//   b200::ConjugateGradient<MatrixType, UpLo, Preconditioner>   drop-in for Eigen::ConjugateGradient
//   b200::BiCGSTAB<MatrixType, Preconditioner>                  drop-in for Eigen::BiCGSTAB
//
// Both derive from Eigen::IterativeSolverBase<Derived> exactly as third-party solvers do
// (unsupported/Eigen/src/IterativeSolvers/GMRES.h:225-327), so compute / analyzePattern / factorize / solve /
// solveWithGuess / setTolerance / setMaxIterations / iterations / error / info, multi-column and sparse right-hand
// sides, Map / uncompressed / expression inputs all behave as in the reference (IterativeSolverBase.h:142-440).
// Only two things are replaced:
//   * compute / analyzePattern / factorize additionally hand the CSR arrays of the grabbed matrix to
//     b200s_analyze_pattern / b200s_factorize_f64 (C ABI, include/b200sparse.h) -- in the manner of
//     KLUSupport/KLUSupport.h:60-115 wrapping klu_*;
//   * _solve_vector_with_guess_impl (ConjugateGradient.h:197-221, BiCGSTAB.h:193-204) calls b200s_*_solve_f64 instead
//     of internal::conjugate_gradient / internal::bicgstab.
// Header-only; needs Eigen on the include path and libb200sparse.so at link time.  There is no CPU fallback: with an
// unsupported instantiation the code does not compile, without a B200 info() reports InvalidInput.
#ifndef B200_ITERATIVE_SOLVERS_H
#define B200_ITERATIVE_SOLVERS_H

#include <Eigen/IterativeLinearSolvers>
#include <Eigen/SparseCore>

#include <string>
#include <vector>

#include "../b200sparse.h"

namespace b200 {
template <typename MatrixType_, int UpLo_ = Eigen::Lower,
          typename Preconditioner_ = Eigen::DiagonalPreconditioner<typename MatrixType_::Scalar> >
class ConjugateGradient;
template <typename MatrixType_, typename Preconditioner_ = Eigen::DiagonalPreconditioner<typename MatrixType_::Scalar> >
class BiCGSTAB;
}  // namespace b200

namespace Eigen {
namespace internal {
template <typename MatrixType_, int UpLo_, typename Preconditioner_>
struct traits<b200::ConjugateGradient<MatrixType_, UpLo_, Preconditioner_> > {
  typedef MatrixType_ MatrixType;
  typedef Preconditioner_ Preconditioner;
};
template <typename MatrixType_, typename Preconditioner_>
struct traits<b200::BiCGSTAB<MatrixType_, Preconditioner_> > {
  typedef MatrixType_ MatrixType;
  typedef Preconditioner_ Preconditioner;
};
}  // namespace internal
}  // namespace Eigen

namespace b200 {
namespace detail {

template <typename P>
struct precond_id {
  enum { supported = 0, value = -1 };
};
template <typename S>
struct precond_id<Eigen::DiagonalPreconditioner<S> > {
  enum { supported = 1, value = B200S_PRECOND_JACOBI };
};
template <>
struct precond_id<Eigen::IdentityPreconditioner> {
  enum { supported = 1, value = B200S_PRECOND_IDENTITY };
};

// Owns one b200s_handle and the host-side staging that turns an Eigen matrix view into the CSR arrays of the C ABI.
class DeviceSolver {
 public:
  DeviceSolver() : m_handle(0) {}
  ~DeviceSolver() {
    if (m_handle) b200s_destroy(m_handle);
  }
  const std::string& lastError() const { return m_error; }

  // `mat` is the solver's grabbed matrix (Ref<const MatrixType>).  `csr_uplo` is the triangle selection expressed for
  // the ROW-major reading of the arrays that are handed over.
  template <typename ActualMatrix>
  bool analyze(const ActualMatrix& mat, int uplo, bool need_transpose) {
    if (!ensure()) return false;
    typedef typename ActualMatrix::Scalar Scalar;
    typedef typename ActualMatrix::StorageIndex StorageIndex;
    const bool row_major = ActualMatrix::IsRowMajor;
    m_use_copy = false;
    if (need_transpose && !row_major) {
      // BiCGSTAB on a column-major matrix: the kernels need rows of A, the arrays hold rows of A^T.  One host-side
      // conversion per compute() (the reference pays a serial scatter product per iteration instead,
      // SparseDenseProduct.h:85-107).
      m_rowmajor_copy = mat;
      m_use_copy = true;
      return push_pattern(m_rowmajor_copy.rows(), m_rowmajor_copy.nonZeros(), m_rowmajor_copy.outerIndexPtr(),
                          m_rowmajor_copy.innerIndexPtr(), m_rowmajor_copy.innerNonZeroPtr(), B200S_BOTH,
                          m_rowmajor_copy.outerIndexPtr()[m_rowmajor_copy.outerSize()]);
    }
    // Column-major arrays read as CSR are the arrays of A^T (ConjugateGradient.h:202-208 uses the same trick): for
    // a self-adjoint operator that is A itself, with Lower and Upper swapping roles.
    int csr_uplo = uplo;
    if (!row_major && uplo != B200S_BOTH) csr_uplo = (uplo == B200S_LOWER) ? B200S_UPPER : B200S_LOWER;
    const StorageIndex* outer = mat.outerIndexPtr();
    const Eigen::Index span = mat.innerNonZeroPtr() ? Eigen::Index(outer[mat.outerSize()]) : Eigen::Index(mat.nonZeros());
    (void)sizeof(Scalar);
    return push_pattern(mat.outerSize(), mat.nonZeros(), outer, mat.innerIndexPtr(), mat.innerNonZeroPtr(), csr_uplo,
                        span);
  }

  template <typename ActualMatrix>
  bool factorize(const ActualMatrix& mat, int precond) {
    if (!m_handle) return false;
    const double* values = m_use_copy ? m_rowmajor_copy.valuePtr() : mat.valuePtr();
    return check(b200s_factorize_f64(m_handle, values, precond));
  }

  bool solve(bool bicg, const double* b, double* x, bool use_guess, double tol, Eigen::Index max_iters,
             Eigen::Index& iters, double& error, Eigen::ComputationInfo& info) {
    if (!m_handle) {
      info = Eigen::InvalidInput;
      return false;
    }
    int64_t it = 0;
    int inf = 0;
    double err = 0;
    int rc = bicg ? b200s_bicgstab_solve_f64(m_handle, b, x, use_guess ? 1 : 0, tol, max_iters, &it, &err, &inf)
                  : b200s_cg_solve_f64(m_handle, b, x, use_guess ? 1 : 0, tol, max_iters, &it, &err, &inf);
    if (!check(rc)) {
      info = Eigen::InvalidInput;
      return false;
    }
    iters = static_cast<Eigen::Index>(it);
    error = err;
    info = static_cast<Eigen::ComputationInfo>(inf);
    return true;
  }

 private:
  bool ensure() {
    if (m_handle) return true;
    int rc = b200s_create(0, &m_handle);
    if (rc != B200S_OK) {
      m_error = b200s_last_error(0);
      m_handle = 0;
      return false;
    }
    return true;
  }
  bool check(int rc) {
    if (rc == B200S_OK) return true;
    m_error = b200s_last_error(m_handle);
    return false;
  }
  template <typename StorageIndex>
  bool push_pattern(Eigen::Index outer_size, Eigen::Index /*nnz*/, const StorageIndex* outer, const StorageIndex* inner,
                    const StorageIndex* inner_nnz, int uplo, Eigen::Index span) {
    // The C ABI speaks int32 (Eigen's default StorageIndex); wider index types are narrowed once per analyzePattern.
    const int32_t *o = 0, *i = 0, *z = 0;
    if (sizeof(StorageIndex) == sizeof(int32_t)) {
      o = reinterpret_cast<const int32_t*>(outer);
      i = reinterpret_cast<const int32_t*>(inner);
      z = reinterpret_cast<const int32_t*>(inner_nnz);
    } else {
      m_outer32.assign(outer, outer + outer_size + 1);
      m_inner32.assign(inner, inner + span);
      o = m_outer32.data();
      i = m_inner32.data();
      if (inner_nnz) {
        m_innernnz32.assign(inner_nnz, inner_nnz + outer_size);
        z = m_innernnz32.data();
      }
    }
    // `span` = one past the last stored slot: for uncompressed matrices the value/index arrays have holes
    return check(b200s_analyze_pattern(m_handle, outer_size, outer_size, span, o, i, z, uplo, 0));
  }

  b200s_handle* m_handle;
  std::string m_error;
  bool m_use_copy;
  Eigen::SparseMatrix<double, Eigen::RowMajor, int> m_rowmajor_copy;
  std::vector<int32_t> m_outer32, m_inner32, m_innernnz32;
};

}  // namespace detail

// ---------------------------------------------------------------------------------------------------------------
template <typename MatrixType_, int UpLo_, typename Preconditioner_>
class ConjugateGradient : public Eigen::IterativeSolverBase<ConjugateGradient<MatrixType_, UpLo_, Preconditioner_> > {
  typedef Eigen::IterativeSolverBase<ConjugateGradient> Base;
  using Base::m_error;
  using Base::m_info;
  using Base::m_isInitialized;
  using Base::m_iterations;
  using Base::matrix;

 public:
  typedef MatrixType_ MatrixType;
  typedef typename MatrixType::Scalar Scalar;
  typedef typename MatrixType::RealScalar RealScalar;
  typedef Preconditioner_ Preconditioner;
  enum { UpLo = UpLo_ };

  EIGEN_STATIC_ASSERT((Eigen::internal::is_same<Scalar, double>::value), THIS_TYPE_IS_NOT_SUPPORTED)
  EIGEN_STATIC_ASSERT(detail::precond_id<Preconditioner>::supported, THIS_TYPE_IS_NOT_SUPPORTED)

  ConjugateGradient() : Base() {}

  // Base(A) would run the base-class compute (IterativeSolverBase.h:181-187): run ours, which also uploads.
  template <typename MatrixDerived>
  explicit ConjugateGradient(const Eigen::EigenBase<MatrixDerived>& A) : Base() {
    compute(A.derived());
  }

  template <typename MatrixDerived>
  ConjugateGradient& analyzePattern(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::analyzePattern(A.derived());
    if (!m_dev.analyze(matrix(), int(UpLo), false)) m_info = Eigen::InvalidInput;
    return *this;
  }
  template <typename MatrixDerived>
  ConjugateGradient& factorize(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::factorize(A.derived());
    if (!m_dev.factorize(matrix(), detail::precond_id<Preconditioner>::value)) m_info = Eigen::InvalidInput;
    return *this;
  }
  template <typename MatrixDerived>
  ConjugateGradient& compute(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::compute(A.derived());
    if (!m_dev.analyze(matrix(), int(UpLo), false) ||
        !m_dev.factorize(matrix(), detail::precond_id<Preconditioner>::value))
      m_info = Eigen::InvalidInput;
    return *this;
  }

  /** \internal replaces ConjugateGradient.h:197-221 */
  template <typename Rhs, typename Dest>
  void _solve_vector_with_guess_impl(const Rhs& b, Dest& x) const {
    Eigen::Matrix<double, Eigen::Dynamic, 1> bb = b, xx = x;  // contiguous staging (b, x may be strided blocks)
    m_iterations = Base::maxIterations();
    m_error = Base::m_tolerance;
    // solve() hands over x = 0 (IterativeSolverBase.h:402): skip the initial A*x0 product then, as r0 = b exactly
    const bool guess = (xx.array() != 0.0).any();
    m_dev.solve(false, bb.data(), xx.data(), guess, Base::m_tolerance, Base::maxIterations(), m_iterations, m_error,
                m_info);
    x = xx;
  }

  const std::string& lastError() const { return m_dev.lastError(); }

 protected:
  mutable detail::DeviceSolver m_dev;
};

// ---------------------------------------------------------------------------------------------------------------
template <typename MatrixType_, typename Preconditioner_>
class BiCGSTAB : public Eigen::IterativeSolverBase<BiCGSTAB<MatrixType_, Preconditioner_> > {
  typedef Eigen::IterativeSolverBase<BiCGSTAB> Base;
  using Base::m_error;
  using Base::m_info;
  using Base::m_isInitialized;
  using Base::m_iterations;
  using Base::matrix;

 public:
  typedef MatrixType_ MatrixType;
  typedef typename MatrixType::Scalar Scalar;
  typedef typename MatrixType::RealScalar RealScalar;
  typedef Preconditioner_ Preconditioner;

  EIGEN_STATIC_ASSERT((Eigen::internal::is_same<Scalar, double>::value), THIS_TYPE_IS_NOT_SUPPORTED)
  EIGEN_STATIC_ASSERT(detail::precond_id<Preconditioner>::supported, THIS_TYPE_IS_NOT_SUPPORTED)

  BiCGSTAB() : Base() {}
  template <typename MatrixDerived>
  explicit BiCGSTAB(const Eigen::EigenBase<MatrixDerived>& A) : Base() {
    compute(A.derived());
  }

  template <typename MatrixDerived>
  BiCGSTAB& analyzePattern(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::analyzePattern(A.derived());
    if (!m_dev.analyze(matrix(), B200S_BOTH, true)) m_info = Eigen::InvalidInput;
    return *this;
  }
  template <typename MatrixDerived>
  BiCGSTAB& factorize(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::factorize(A.derived());
    // a column-major input is converted to rows at analyze time; refresh that copy's values too
    if (!MatrixType::IsRowMajor && !m_dev.analyze(matrix(), B200S_BOTH, true)) m_info = Eigen::InvalidInput;
    if (!m_dev.factorize(matrix(), detail::precond_id<Preconditioner>::value)) m_info = Eigen::InvalidInput;
    return *this;
  }
  template <typename MatrixDerived>
  BiCGSTAB& compute(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::compute(A.derived());
    if (!m_dev.analyze(matrix(), B200S_BOTH, true) ||
        !m_dev.factorize(matrix(), detail::precond_id<Preconditioner>::value))
      m_info = Eigen::InvalidInput;
    return *this;
  }

  /** \internal replaces BiCGSTAB.h:193-204 */
  template <typename Rhs, typename Dest>
  void _solve_vector_with_guess_impl(const Rhs& b, Dest& x) const {
    Eigen::Matrix<double, Eigen::Dynamic, 1> bb = b, xx = x;
    m_iterations = Base::maxIterations();
    m_error = Base::m_tolerance;
    // solve() hands over x = 0 (IterativeSolverBase.h:402): skip the initial A*x0 product then, as r0 = b exactly
    const bool guess = (xx.array() != 0.0).any();
    m_dev.solve(true, bb.data(), xx.data(), guess, Base::m_tolerance, Base::maxIterations(), m_iterations, m_error,
                m_info);
    x = xx;
  }

  const std::string& lastError() const { return m_dev.lastError(); }

 protected:
  mutable detail::DeviceSolver m_dev;
};

}  // namespace b200

#endif  // B200_ITERATIVE_SOLVERS_H

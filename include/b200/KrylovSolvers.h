// b200/KrylovSolvers.h -- reference-side binding of the solvers that reuse the hot path's device primitives
// (SURVEY 8f rank 3):
//
//   b200::LeastSquaresConjugateGradient<MatrixType, Preconditioner>   Eigen::LeastSquaresConjugateGradient
//                                              (Eigen/src/IterativeLinearSolvers/LeastSquareConjugateGradient.h:95-213)
//   b200::MINRES<MatrixType, UpLo, Preconditioner>                    Eigen::MINRES (unsupported/.../MINRES.h:142-262)
//   b200::GMRES<MatrixType, Preconditioner>                           Eigen::GMRES  (unsupported/.../GMRES.h:217-338)
//
// Same construction as b200/IterativeSolvers.h: each class derives from Eigen::IterativeSolverBase<Derived>, hands the
// grabbed matrix to the C ABI in compute / analyzePattern / factorize and replaces _solve_vector_with_guess_impl by
// one call of b200s_lscg_solve_f64 / b200s_minres_solve_f64 / b200s_gmres_solve_f64.  Scalar = double, one GPU.
// Preconditioners: LeastSquareDiagonalPreconditioner / IdentityPreconditioner for LSCG, DiagonalPreconditioner /
// IdentityPreconditioner for MINRES and GMRES; anything else does not compile (no CPU fallback).
#ifndef B200_KRYLOV_SOLVERS_H
#define B200_KRYLOV_SOLVERS_H

#include "IterativeSolvers.h"

namespace b200 {
template <typename MatrixType_, typename Preconditioner_ = Eigen::LeastSquareDiagonalPreconditioner<typename MatrixType_::Scalar> >
class LeastSquaresConjugateGradient;
template <typename MatrixType_, int UpLo_ = Eigen::Lower, typename Preconditioner_ = Eigen::IdentityPreconditioner>
class MINRES;
template <typename MatrixType_, typename Preconditioner_ = Eigen::DiagonalPreconditioner<typename MatrixType_::Scalar> >
class GMRES;
}  // namespace b200

namespace Eigen {
namespace internal {
template <typename MatrixType_, typename Preconditioner_>
struct traits<b200::LeastSquaresConjugateGradient<MatrixType_, Preconditioner_> > {
  typedef MatrixType_ MatrixType;
  typedef Preconditioner_ Preconditioner;
};
template <typename MatrixType_, int UpLo_, typename Preconditioner_>
struct traits<b200::MINRES<MatrixType_, UpLo_, Preconditioner_> > {
  typedef MatrixType_ MatrixType;
  typedef Preconditioner_ Preconditioner;
};
template <typename MatrixType_, typename Preconditioner_>
struct traits<b200::GMRES<MatrixType_, Preconditioner_> > {
  typedef MatrixType_ MatrixType;
  typedef Preconditioner_ Preconditioner;
};
}  // namespace internal
}  // namespace Eigen

namespace b200 {
namespace detail {
template <typename S>
struct precond_id<Eigen::LeastSquareDiagonalPreconditioner<S> > {
  enum { supported = 1, value = B200S_PRECOND_JACOBI };
};
}  // namespace detail

// ---------------------------------------------------------------------------------------------------------------
template <typename MatrixType_, typename Preconditioner_>
class LeastSquaresConjugateGradient
    : public Eigen::IterativeSolverBase<LeastSquaresConjugateGradient<MatrixType_, Preconditioner_> > {
  typedef Eigen::IterativeSolverBase<LeastSquaresConjugateGradient> Base;
  using Base::m_error;
  using Base::m_info;
  using Base::m_iterations;
  using Base::matrix;

 public:
  typedef MatrixType_ MatrixType;
  typedef typename MatrixType::Scalar Scalar;
  typedef typename MatrixType::RealScalar RealScalar;
  typedef Preconditioner_ Preconditioner;
  EIGEN_STATIC_ASSERT((Eigen::internal::is_same<Scalar, double>::value), THIS_TYPE_IS_NOT_SUPPORTED)
  EIGEN_STATIC_ASSERT(detail::precond_id<Preconditioner>::supported, THIS_TYPE_IS_NOT_SUPPORTED)

  LeastSquaresConjugateGradient() : Base() {}
  template <typename MatrixDerived>
  explicit LeastSquaresConjugateGradient(const Eigen::EigenBase<MatrixDerived>& A) : Base() {
    compute(A.derived());
  }

  template <typename MatrixDerived>
  LeastSquaresConjugateGradient& analyzePattern(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::analyzePattern(A.derived());
    if (!upload(true, false)) m_info = Eigen::InvalidInput;
    return *this;
  }
  template <typename MatrixDerived>
  LeastSquaresConjugateGradient& factorize(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::factorize(A.derived());
    if (!upload(!MatrixType::IsRowMajor, true)) m_info = Eigen::InvalidInput;  // host-side conversions hold values too
    return *this;
  }
  template <typename MatrixDerived>
  LeastSquaresConjugateGradient& compute(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::compute(A.derived());
    if (!upload(true, true)) m_info = Eigen::InvalidInput;
    return *this;
  }

  /** \internal replaces LeastSquareConjugateGradient.h:197-208 */
  template <typename Rhs, typename Dest>
  void _solve_vector_with_guess_impl(const Rhs& b, Dest& x) const {
    typedef Eigen::Matrix<double, Eigen::Dynamic, 1> Vec;
    m_iterations = Base::maxIterations();
    m_error = Base::m_tolerance;
    Vec bb = b, xx = x;
    const bool guess = (xx.array() != 0.0).any();
    int64_t it = 0;
    double err = 0;
    int info = 0;
    if (!m_dev.handle() || !m_dev_t.handle() ||
        b200s_lscg_solve_f64(m_dev.handle(), m_dev_t.handle(), bb.data(), xx.data(), guess ? 1 : 0, m_error,
                             Base::maxIterations(), detail::precond_id<Preconditioner>::value,
                             MatrixType::IsRowMajor ? 0 : 1, &it, &err, &info) != B200S_OK) {
      m_info = Eigen::InvalidInput;
      return;
    }
    x = xx;
    m_iterations = static_cast<Eigen::Index>(it);
    m_error = err;
    m_info = static_cast<Eigen::ComputationInfo>(info);
  }

 private:
  // A (rows of A) into m_dev, A^T into m_dev_t.  Column-major arrays read as CSR are already A^T.
  bool upload(bool pattern, bool values) {
    typedef Eigen::SparseMatrix<double, Eigen::RowMajor, int> Csr;
    if (MatrixType::IsRowMajor) {
      if (pattern) m_t = Csr(matrix().transpose());
      else m_t = Csr(matrix().transpose());
      m_t.makeCompressed();
    }
    bool ok = true;
    if (pattern) {
      ok = ok && m_dev.analyze(matrix(), B200S_BOTH, true);  // rows of A (a column-major A is converted once)
      ok = ok && (MatrixType::IsRowMajor ? m_dev_t.analyze(m_t, B200S_BOTH, false) : m_dev_t.analyze(matrix(), B200S_BOTH, false));
    }
    if (values) {
      ok = ok && m_dev.factorize(matrix(), B200S_PRECOND_IDENTITY);
      ok = ok && (MatrixType::IsRowMajor ? m_dev_t.factorize(m_t, B200S_PRECOND_IDENTITY) : m_dev_t.factorize(matrix(), B200S_PRECOND_IDENTITY));
    }
    return ok;
  }
  mutable detail::DeviceSolver<double> m_dev, m_dev_t;
  Eigen::SparseMatrix<double, Eigen::RowMajor, int> m_t;
};

// ---------------------------------------------------------------------------------------------------------------
template <typename MatrixType_, int UpLo_, typename Preconditioner_>
class MINRES : public Eigen::IterativeSolverBase<MINRES<MatrixType_, UpLo_, Preconditioner_> > {
  typedef Eigen::IterativeSolverBase<MINRES> Base;
  using Base::m_error;
  using Base::m_info;
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

  MINRES() : Base() {}
  template <typename MatrixDerived>
  explicit MINRES(const Eigen::EigenBase<MatrixDerived>& A) : Base() {
    compute(A.derived());
  }
  template <typename MatrixDerived>
  MINRES& analyzePattern(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::analyzePattern(A.derived());
    if (!m_dev.analyze(matrix(), int(UpLo), false)) m_info = Eigen::InvalidInput;
    return *this;
  }
  template <typename MatrixDerived>
  MINRES& factorize(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::factorize(A.derived());
    if (!m_dev.factorize(matrix(), detail::precond_id<Preconditioner>::value, Base::m_preconditioner)) m_info = Eigen::InvalidInput;
    return *this;
  }
  template <typename MatrixDerived>
  MINRES& compute(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::compute(A.derived());
    if (!m_dev.analyze(matrix(), int(UpLo), false) || !m_dev.factorize(matrix(), detail::precond_id<Preconditioner>::value, Base::m_preconditioner))
      m_info = Eigen::InvalidInput;
    return *this;
  }
  /** \internal replaces MINRES.h:236-262 */
  template <typename Rhs, typename Dest>
  void _solve_vector_with_guess_impl(const Rhs& b, Dest& x) const {
    typedef Eigen::Matrix<double, Eigen::Dynamic, 1> Vec;
    m_iterations = Base::maxIterations();
    m_error = Base::m_tolerance;
    Vec bb = b, xx = x;
    const bool guess = (xx.array() != 0.0).any();
    int64_t it = 0;
    double err = 0;
    int info = 0;
    if (!m_dev.handle() || b200s_minres_solve_f64(m_dev.handle(), bb.data(), xx.data(), guess ? 1 : 0, m_error,
                                                  Base::maxIterations(), &it, &err, &info) != B200S_OK) {
      m_info = Eigen::InvalidInput;
      return;
    }
    x = xx;
    m_iterations = static_cast<Eigen::Index>(it);
    m_error = err;
    m_info = static_cast<Eigen::ComputationInfo>(info);
  }

 protected:
  mutable detail::DeviceSolver<double> m_dev;
};

// ---------------------------------------------------------------------------------------------------------------
template <typename MatrixType_, typename Preconditioner_>
class GMRES : public Eigen::IterativeSolverBase<GMRES<MatrixType_, Preconditioner_> > {
  typedef Eigen::IterativeSolverBase<GMRES> Base;
  using Base::m_error;
  using Base::m_info;
  using Base::m_iterations;
  using Base::matrix;

 public:
  typedef MatrixType_ MatrixType;
  typedef typename MatrixType::Scalar Scalar;
  typedef typename MatrixType::RealScalar RealScalar;
  typedef Preconditioner_ Preconditioner;
  EIGEN_STATIC_ASSERT((Eigen::internal::is_same<Scalar, double>::value), THIS_TYPE_IS_NOT_SUPPORTED)
  EIGEN_STATIC_ASSERT(detail::precond_id<Preconditioner>::supported, THIS_TYPE_IS_NOT_SUPPORTED)

  GMRES() : Base(), m_restart(30) {}
  template <typename MatrixDerived>
  explicit GMRES(const Eigen::EigenBase<MatrixDerived>& A) : Base(), m_restart(30) {
    compute(A.derived());
  }
  Eigen::Index get_restart() { return m_restart; }
  void set_restart(const Eigen::Index restart) { m_restart = restart; }

  template <typename MatrixDerived>
  GMRES& analyzePattern(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::analyzePattern(A.derived());
    if (!m_dev.analyze(matrix(), B200S_BOTH, true)) m_info = Eigen::InvalidInput;
    return *this;
  }
  template <typename MatrixDerived>
  GMRES& factorize(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::factorize(A.derived());
    if (!MatrixType::IsRowMajor && !m_dev.analyze(matrix(), B200S_BOTH, true)) m_info = Eigen::InvalidInput;
    if (!m_dev.factorize(matrix(), detail::precond_id<Preconditioner>::value, Base::m_preconditioner)) m_info = Eigen::InvalidInput;
    return *this;
  }
  template <typename MatrixDerived>
  GMRES& compute(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::compute(A.derived());
    if (!m_dev.analyze(matrix(), B200S_BOTH, true) || !m_dev.factorize(matrix(), detail::precond_id<Preconditioner>::value, Base::m_preconditioner))
      m_info = Eigen::InvalidInput;
    return *this;
  }
  /** \internal replaces GMRES.h:317-325 */
  template <typename Rhs, typename Dest>
  void _solve_vector_with_guess_impl(const Rhs& b, Dest& x) const {
    typedef Eigen::Matrix<double, Eigen::Dynamic, 1> Vec;
    m_iterations = Base::maxIterations();
    m_error = Base::m_tolerance;
    Vec bb = b, xx = x;
    const bool guess = (xx.array() != 0.0).any();
    int64_t it = 0;
    double err = 0;
    int info = 0;
    if (!m_dev.handle() || b200s_gmres_solve_f64(m_dev.handle(), bb.data(), xx.data(), guess ? 1 : 0, m_error,
                                                 Base::maxIterations(), m_restart, &it, &err, &info) != B200S_OK) {
      m_info = Eigen::InvalidInput;
      return;
    }
    x = xx;
    m_iterations = static_cast<Eigen::Index>(it);
    m_error = err;
    m_info = static_cast<Eigen::ComputationInfo>(info);
  }

 protected:
  mutable detail::DeviceSolver<double> m_dev;
  Eigen::Index m_restart;
};

}  // namespace b200

#endif  // B200_KRYLOV_SOLVERS_H

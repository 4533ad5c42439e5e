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
// Scalar = double or float (the reference's real instantiations, ConjugateGradient.h:157-160); complex does not compile.
// b200/SparseOperator.h holds the SpMV-only drop-in (operator concept, doc/examples/matrixfree_cg.cpp:14-76).
//
// Row-partitioned use from C++ (one process per GPU): call setDistributed(rank, world, row_starts, allgather, ctx)
// before compute(); compute() then takes THIS rank's row block as a rows_local x N row-major matrix with global column
// indices, solve() takes this rank's block of b and returns an N-vector whose segment [row_starts[rank],
// row_starts[rank+1]) holds this rank's block of x (the rest is zero).  `allgather` is only used during setup and for
// the status exchange before a launch (MPI_Allgather, or any bootstrap the host has).
// Header-only; needs Eigen on the include path and libb200sparse.so at link time.  There is no CPU fallback: with an
// unsupported instantiation the code does not compile, without a B200 info() reports InvalidInput.
#ifndef B200_ITERATIVE_SOLVERS_H
#define B200_ITERATIVE_SOLVERS_H

#include <Eigen/IterativeLinearSolvers>
#include <Eigen/SparseCore>

#include <cstring>
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
// Incomplete factorizations (IncompleteLUT.h, IncompleteCholesky.h): the preconditioner object the solver owns computes
// its factor on the host exactly as in the reference (IterativeSolverBase::compute calls m_preconditioner.compute,
// IterativeSolverBase.h:196-247); the factor is then handed to the device, which applies it in every iteration
// (b200s_set_preconditioner: level-scheduled triangular solves).  double only.
template <typename Idx>
struct precond_id<Eigen::IncompleteLUT<double, Idx> > {
  enum { supported = 1, value = B200S_PRECOND_FACTORS };
};
template <int UpLo, typename Ordering>
struct precond_id<Eigen::IncompleteCholesky<double, UpLo, Ordering> > {
  enum { supported = 1, value = B200S_PRECOND_FACTORS };
};

// factors_of<P>::make: the b200s_factors object for what a preconditioner of the reference holds (0 = nothing to hand over)
template <typename P>
struct factors_of {
  enum { has = 0 };
  static b200s_factors* make(const P&) { return 0; }
};
template <typename Idx>
struct factors_of<Eigen::IncompleteLUT<double, Idx> > {
  enum { has = 1 };
  typedef Eigen::IncompleteLUT<double, Idx> Pre;
  // m_lu / m_P are protected (IncompleteLUT.h:183-189): pointers to members named through a derived class reach them
  struct Peek : Pre {
    static const typename Pre::FactorType& lu(const Pre& p) { return p.*(&Peek::m_lu); }
    static const Eigen::PermutationMatrix<Eigen::Dynamic, Eigen::Dynamic, Idx>& perm(const Pre& p) { return p.*(&Peek::m_P); }
  };
  static b200s_factors* make(const Pre& pre) {
    const typename Pre::FactorType& lu = Peek::lu(pre);
    const Eigen::Index n = lu.rows(), nz = lu.outerIndexPtr()[n];
    std::vector<int32_t> rp(lu.outerIndexPtr(), lu.outerIndexPtr() + n + 1), ci(lu.innerIndexPtr(), lu.innerIndexPtr() + nz);
    const Idx* pi = Peek::perm(pre).indices().data();
    std::vector<int32_t> perm(pi, pi + n);
    b200s_factors* f = 0;
    b200s_factors_from_ilut_f64(n, rp.data(), ci.data(), lu.valuePtr(), perm.data(), &f);
    return f;
  }
};
template <int UpLo, typename Ordering>
struct factors_of<Eigen::IncompleteCholesky<double, UpLo, Ordering> > {
  enum { has = 1 };
  typedef Eigen::IncompleteCholesky<double, UpLo, Ordering> Pre;
  static b200s_factors* make(const Pre& pre) {
    const typename Pre::FactorType& L = pre.matrixL();
    const Eigen::Index n = L.cols(), nz = L.outerIndexPtr()[n];
    std::vector<int32_t> cp(L.outerIndexPtr(), L.outerIndexPtr() + n + 1), ri(L.innerIndexPtr(), L.innerIndexPtr() + nz);
    const Eigen::Index ps = pre.permutationP().size();  // 0 for NaturalOrdering (IncompleteCholesky.h:100-102)
    std::vector<int32_t> perm(pre.permutationP().indices().data(), pre.permutationP().indices().data() + ps);
    b200s_factors* f = 0;
    b200s_factors_from_ichol_f64(n, cp.data(), ri.data(), L.valuePtr(), pre.scalingS().data(), ps ? perm.data() : 0, &f);
    return f;
  }
};

// The C ABI entry points of one scalar type.
template <typename S>
struct abi {
  enum { supported = 0 };
};
template <>
struct abi<double> {
  enum { supported = 1, has_multi = 1 };
  static int solve_multi(b200s_handle* h, int64_t nc, const double* B, int64_t ldb, double* X, int64_t ldx, int g,
                         double tol, int64_t mi, int64_t* it, double* err, int* info) {
    return b200s_cg_solve_multi_f64(h, nc, B, ldb, X, ldx, g, tol, mi, it, err, info);
  }
  static int factorize(b200s_handle* h, const double* v, int p) { return b200s_factorize_f64(h, v, p); }
  static int spmv(b200s_handle* h, const double* x, double* y) { return b200s_spmv_f64(h, x, y); }
  static int solve(b200s_handle* h, bool bicg, const double* b, double* x, int g, double tol, int64_t mi, int64_t* it,
                   double* err, int* info) {
    return bicg ? b200s_bicgstab_solve_f64(h, b, x, g, tol, mi, it, err, info)
                : b200s_cg_solve_f64(h, b, x, g, tol, mi, it, err, info);
  }
};
template <>
struct abi<float> {
  enum { supported = 1, has_multi = 0 };
  static int solve_multi(b200s_handle*, int64_t, const float*, int64_t, float*, int64_t, int, double, int64_t, int64_t*,
                         double*, int*) {
    return B200S_ERR_UNSUPPORTED;
  }
  static int factorize(b200s_handle* h, const float* v, int p) { return b200s_factorize_f32(h, v, p); }
  static int spmv(b200s_handle* h, const float* x, float* y) { return b200s_spmv_f32(h, x, y); }
  static int solve(b200s_handle* h, bool bicg, const float* b, float* x, int g, double tol, int64_t mi, int64_t* it,
                   double* err, int* info) {
    return bicg ? b200s_bicgstab_solve_f32(h, b, x, g, tol, mi, it, err, info)
                : b200s_cg_solve_f32(h, b, x, g, tol, mi, it, err, info);
  }
};

// Owns one b200s_handle and the host-side staging that turns an Eigen matrix view into the CSR arrays of the C ABI.
template <typename Scalar>
class DeviceSolver {
 public:
  DeviceSolver() : m_handle(0), m_use_copy(false), m_rows(0), m_cols(0) {
    std::memset(&m_cfg, 0, sizeof(m_cfg));
    m_cfg.struct_size = sizeof(m_cfg);
    m_cfg.device = -1;
    m_cfg.world = 1;
  }
  ~DeviceSolver() { reset(); }
  const std::string& lastError() const { return m_error; }
  b200s_handle* handle() const { return m_handle; }
  Eigen::Index rows() const { return m_rows; }
  Eigen::Index cols() const { return m_cols; }
  int world() const { return m_cfg.world; }
  Eigen::Index rowStart() const { return m_cfg.world > 1 ? Eigen::Index(m_row_starts[m_cfg.rank]) : 0; }

  // Must precede analyze(); a new configuration drops the device state.
  void configure(const b200s_config& cfg, const int64_t* row_starts) {
    reset();
    m_cfg = cfg;
    m_cfg.struct_size = sizeof(m_cfg);
    if (m_cfg.world <= 0) m_cfg.world = 1;
    m_row_starts.clear();
    if (row_starts) m_row_starts.assign(row_starts, row_starts + m_cfg.world + 1);
  }

  // `mat` is the solver's grabbed matrix (Ref<const MatrixType>).  `csr_uplo` is the triangle selection expressed for
  // the ROW-major reading of the arrays that are handed over.
  template <typename ActualMatrix>
  bool analyze(const ActualMatrix& mat, int uplo, bool need_transpose) {
    if (!ensure()) return false;
    typedef typename ActualMatrix::StorageIndex StorageIndex;
    const bool row_major = ActualMatrix::IsRowMajor;
    m_use_copy = false;
    if (need_transpose && !row_major) {
      // A general operator on a column-major matrix: the kernels need rows of A, the arrays hold rows of A^T.  One
      // host-side conversion per compute() (the reference pays a serial scatter product per iteration instead,
      // SparseDenseProduct.h:85-107).
      m_rowmajor_copy = mat;
      m_use_copy = true;
      return push_pattern(m_rowmajor_copy.rows(), m_rowmajor_copy.cols(), m_rowmajor_copy.outerIndexPtr(),
                          m_rowmajor_copy.innerIndexPtr(), m_rowmajor_copy.innerNonZeroPtr(), B200S_BOTH,
                          m_rowmajor_copy.outerIndexPtr()[m_rowmajor_copy.outerSize()]);
    }
    // Column-major arrays read as CSR are the arrays of A^T (ConjugateGradient.h:202-208 uses the same trick): for
    // a self-adjoint operator that is A itself, with Lower and Upper swapping roles.
    int csr_uplo = uplo;
    if (!row_major && uplo != B200S_BOTH) csr_uplo = (uplo == B200S_LOWER) ? B200S_UPPER : B200S_LOWER;
    const StorageIndex* outer = mat.outerIndexPtr();
    // one past the last slot the outer index references: equals nonZeros() for a compressed matrix that starts at
    // slot 0, and stays correct for uncompressed storage (holes) and for a Map/Ref of an inner panel (outer[0] > 0)
    const Eigen::Index span = Eigen::Index(outer[mat.outerSize()]);
    return push_pattern(mat.outerSize(), mat.innerSize(), outer, mat.innerIndexPtr(), mat.innerNonZeroPtr(), csr_uplo,
                        span);
  }

  template <typename ActualMatrix>
  bool factorize(const ActualMatrix& mat, int precond) {
    if (!m_handle) return false;
    const Scalar* values = m_use_copy ? m_rowmajor_copy.valuePtr() : mat.valuePtr();
    return check(abi<Scalar>::factorize(m_handle, values, precond));
  }

  // As above, then the preconditioner object's factor goes to the device (nothing to do for Jacobi / identity).
  template <typename ActualMatrix, typename P>
  bool factorize(const ActualMatrix& mat, int precond, const P& pre) {
    return factorize_with(mat, precond, pre, Eigen::internal::bool_constant<bool(factors_of<P>::has)>());
  }

  bool solve(bool bicg, const Scalar* b, Scalar* x, bool use_guess, double tol, Eigen::Index max_iters,
             Eigen::Index& iters, double& error, Eigen::ComputationInfo& info) {
    if (!m_handle) {
      info = Eigen::InvalidInput;
      return false;
    }
    int64_t it = 0;
    int inf = 0;
    double err = 0;
    if (!check(abi<Scalar>::solve(m_handle, bicg, b, x, use_guess ? 1 : 0, tol, max_iters, &it, &err, &inf))) {
      info = Eigen::InvalidInput;
      return false;
    }
    iters = static_cast<Eigen::Index>(it);
    error = err;
    info = static_cast<Eigen::ComputationInfo>(inf);
    return true;
  }

  // CG for all columns of a column-major block at once (b200s_cg_solve_multi_f64)
  bool solve_multi(Eigen::Index ncols, const Scalar* B, Eigen::Index ldb, Scalar* X, Eigen::Index ldx, bool use_guess,
                   double tol, Eigen::Index max_iters, std::vector<int64_t>& iters, std::vector<double>& errors,
                   std::vector<int>& infos) {
    if (!m_handle) return false;
    iters.assign(ncols, 0);
    errors.assign(ncols, 0.0);
    infos.assign(ncols, 0);
    return check(abi<Scalar>::solve_multi(m_handle, ncols, B, ldb, X, ldx, use_guess ? 1 : 0, tol, max_iters,
                                          iters.data(), errors.data(), infos.data()));
  }

  // y = A x through b200s_spmv_* (x: cols() entries on one GPU, this rank's rows otherwise; y: this rank's rows)
  bool multiply(const Scalar* x, Scalar* y) {
    if (!m_handle) return false;
    return check(abi<Scalar>::spmv(m_handle, x, y));
  }

 private:
  template <typename ActualMatrix, typename P>
  bool factorize_with(const ActualMatrix& mat, int precond, const P&, Eigen::internal::false_type) {
    return factorize(mat, precond);
  }
  template <typename ActualMatrix, typename P>
  bool factorize_with(const ActualMatrix& mat, int, const P& pre, Eigen::internal::true_type) {
    if (!factorize(mat, B200S_PRECOND_JACOBI)) return false;
    if (pre.info() != Eigen::Success) return true;  // the solver's info() already reports the failed factorization
    b200s_factors* f = factors_of<P>::make(pre);
    if (!f) {
      m_error = b200s_last_error(0);
      return false;
    }
    const bool ok = check(b200s_set_preconditioner(m_handle, f));
    b200s_factors_destroy(f);
    return ok;
  }
  void reset() {
    if (m_handle) b200s_destroy(m_handle);
    m_handle = 0;
  }
  bool ensure() {
    if (m_handle) return true;
    int rc = b200s_create(&m_cfg, &m_handle);
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
  bool push_pattern(Eigen::Index outer_size, Eigen::Index inner_size, const StorageIndex* outer,
                    const StorageIndex* inner, const StorageIndex* inner_nnz, int uplo, Eigen::Index span) {
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
    m_rows = outer_size;
    m_cols = inner_size;
    // `span` = one past the last stored slot: for uncompressed matrices the value/index arrays have holes
    return check(b200s_analyze_pattern(m_handle, outer_size, inner_size, span, o, i, z, uplo,
                                       m_cfg.world > 1 ? m_row_starts.data() : 0));
  }

  b200s_handle* m_handle;
  b200s_config m_cfg;
  std::vector<int64_t> m_row_starts;
  std::string m_error;
  bool m_use_copy;
  Eigen::Index m_rows, m_cols;
  Eigen::SparseMatrix<Scalar, Eigen::RowMajor, int> m_rowmajor_copy;
  std::vector<int32_t> m_outer32, m_inner32, m_innernnz32;
};

// Shared by both solver classes: contiguous staging of b / x, the single-GPU and the row-partitioned calling shapes.
template <typename Scalar, typename Rhs, typename Dest>
void solve_vector(DeviceSolver<Scalar>& dev, bool bicg, const Rhs& b, Dest& x, double tol, Eigen::Index max_iters,
                  Eigen::Index& iters, double& error, Eigen::ComputationInfo& info) {
  typedef Eigen::Matrix<Scalar, Eigen::Dynamic, 1> Vec;
  Vec bb = b;  // contiguous staging (b, x may be strided blocks or expressions)
  if (dev.world() > 1) {
    // x is an N-vector (the base class sizes it by cols()); this rank works on its own segment
    Vec xx = x.segment(dev.rowStart(), dev.rows());
    const bool guess = (xx.array() != Scalar(0)).any();
    dev.solve(bicg, bb.data(), xx.data(), guess, tol, max_iters, iters, error, info);
    x.setZero();
    x.segment(dev.rowStart(), dev.rows()) = xx;
    return;
  }
  Vec xx = x;
  // solve() hands over x = 0 (IterativeSolverBase.h:402): skip the initial A*x0 product then, as r0 = b exactly
  const bool guess = (xx.array() != Scalar(0)).any();
  dev.solve(bicg, bb.data(), xx.data(), guess, tol, max_iters, iters, error, info);
  x = xx;
}

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

  EIGEN_STATIC_ASSERT(detail::abi<Scalar>::supported, THIS_TYPE_IS_NOT_SUPPORTED)
  EIGEN_STATIC_ASSERT(detail::precond_id<Preconditioner>::supported, THIS_TYPE_IS_NOT_SUPPORTED)

  ConjugateGradient() : Base() {}

  // Base(A) would run the base-class compute (IterativeSolverBase.h:181-187): run ours, which also uploads.
  template <typename MatrixDerived>
  explicit ConjugateGradient(const Eigen::EigenBase<MatrixDerived>& A) : Base() {
    compute(A.derived());
  }

  /** Row-partitioned run, one process per GPU (see the file header).  Call before compute(). */
  ConjugateGradient& setDistributed(int rank, int world, const int64_t* row_starts, b200s_allgather_fn allgather,
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

  template <typename MatrixDerived>
  ConjugateGradient& analyzePattern(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::analyzePattern(A.derived());
    if (!m_dev.analyze(matrix(), int(UpLo), false)) m_info = Eigen::InvalidInput;
    return *this;
  }
  template <typename MatrixDerived>
  ConjugateGradient& factorize(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::factorize(A.derived());
    if (!m_dev.factorize(matrix(), detail::precond_id<Preconditioner>::value, Base::m_preconditioner)) m_info = Eigen::InvalidInput;
    return *this;
  }
  template <typename MatrixDerived>
  ConjugateGradient& compute(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::compute(A.derived());
    if (!m_dev.analyze(matrix(), int(UpLo), false) ||
        !m_dev.factorize(matrix(), detail::precond_id<Preconditioner>::value, Base::m_preconditioner))
      m_info = Eigen::InvalidInput;
    return *this;
  }

  /** \internal replaces ConjugateGradient.h:197-221 */
  template <typename Rhs, typename Dest>
  void _solve_vector_with_guess_impl(const Rhs& b, Dest& x) const {
    m_iterations = Base::maxIterations();
    m_error = Base::m_tolerance;
    double err = static_cast<double>(m_error);
    detail::solve_vector<Scalar>(m_dev, false, b, x, static_cast<double>(Base::m_tolerance), Base::maxIterations(),
                                 m_iterations, err, m_info);
    m_error = static_cast<RealScalar>(err);
  }

  using Base::_solve_with_guess_impl;
  /** \internal replaces the per-column loop of IterativeSolverBase.h:366-389 for dense multi-column right-hand sides:
   * all columns share one stream of the matrix per iteration; every column's x, and the iterations() / error() /
   * info() the loop would leave behind (last column / worst column), are those of the sequential loop. */
  template <typename Rhs, typename DestDerived>
  typename Eigen::internal::enable_if<Rhs::ColsAtCompileTime != 1 && DestDerived::ColsAtCompileTime != 1>::type
  _solve_with_guess_impl(const Rhs& b, Eigen::MatrixBase<DestDerived>& aDest) const {
    if (!detail::abi<Scalar>::has_multi || m_dev.world() > 1 || b.cols() < 2) {
      Base::_solve_with_guess_impl(b, aDest);
      return;
    }
    eigen_assert(Base::rows() == b.rows());
    typedef Eigen::Matrix<Scalar, Eigen::Dynamic, Eigen::Dynamic> Dense;  // column-major, contiguous
    Dense B = b, X = aDest.derived();
    const bool guess = (X.array() != Scalar(0)).any();
    std::vector<int64_t> its;
    std::vector<double> errs;
    std::vector<int> infos;
    m_iterations = Base::maxIterations();
    m_error = Base::m_tolerance;
    if (!m_dev.solve_multi(B.cols(), B.data(), B.outerStride(), X.data(), X.outerStride(), guess,
                           static_cast<double>(Base::m_tolerance), Base::maxIterations(), its, errs, infos)) {
      m_info = Eigen::InvalidInput;
      return;
    }
    aDest.derived() = X;
    Eigen::ComputationInfo global_info = Eigen::Success;
    for (std::size_t k = 0; k < infos.size(); ++k) {  // IterativeSolverBase.h:383-386
      if (infos[k] == Eigen::NumericalIssue) global_info = Eigen::NumericalIssue;
      else if (infos[k] == Eigen::NoConvergence) global_info = Eigen::NoConvergence;
    }
    m_iterations = static_cast<Eigen::Index>(its.back());
    m_error = static_cast<RealScalar>(errs.back());
    m_info = global_info;
  }

  const std::string& lastError() const { return m_dev.lastError(); }
  b200s_handle* handle() const { return m_dev.handle(); }

 protected:
  mutable detail::DeviceSolver<Scalar> m_dev;
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

  EIGEN_STATIC_ASSERT(detail::abi<Scalar>::supported, THIS_TYPE_IS_NOT_SUPPORTED)
  EIGEN_STATIC_ASSERT(detail::precond_id<Preconditioner>::supported, THIS_TYPE_IS_NOT_SUPPORTED)

  BiCGSTAB() : Base() {}
  template <typename MatrixDerived>
  explicit BiCGSTAB(const Eigen::EigenBase<MatrixDerived>& A) : Base() {
    compute(A.derived());
  }

  /** Row-partitioned run, one process per GPU (see the file header).  Call before compute(). */
  BiCGSTAB& setDistributed(int rank, int world, const int64_t* row_starts, b200s_allgather_fn allgather,
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
    if (!m_dev.factorize(matrix(), detail::precond_id<Preconditioner>::value, Base::m_preconditioner)) m_info = Eigen::InvalidInput;
    return *this;
  }
  template <typename MatrixDerived>
  BiCGSTAB& compute(const Eigen::EigenBase<MatrixDerived>& A) {
    Base::compute(A.derived());
    if (!m_dev.analyze(matrix(), B200S_BOTH, true) ||
        !m_dev.factorize(matrix(), detail::precond_id<Preconditioner>::value, Base::m_preconditioner))
      m_info = Eigen::InvalidInput;
    return *this;
  }

  /** \internal replaces BiCGSTAB.h:193-204 */
  template <typename Rhs, typename Dest>
  void _solve_vector_with_guess_impl(const Rhs& b, Dest& x) const {
    m_iterations = Base::maxIterations();
    m_error = Base::m_tolerance;
    double err = static_cast<double>(m_error);
    detail::solve_vector<Scalar>(m_dev, true, b, x, static_cast<double>(Base::m_tolerance), Base::maxIterations(),
                                 m_iterations, err, m_info);
    m_error = static_cast<RealScalar>(err);
  }

  const std::string& lastError() const { return m_dev.lastError(); }
  b200s_handle* handle() const { return m_dev.handle(); }

 protected:
  mutable detail::DeviceSolver<Scalar> m_dev;
};

}  // namespace b200

#endif  // B200_ITERATIVE_SOLVERS_H

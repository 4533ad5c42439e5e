// oracle/ref_eigen.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin extern "C" wrapper around the UNMODIFIED reference (Eigen headers where they lie under
// /root/reference), compiled by oracle/Makefile into oracle/_ref/libeigen_ref.so.  It contains no
// algorithm of its own: every function instantiates the reference's own templates
//   SparseMatrix<T,RowMajor,int> * Vector   Eigen/src/SparseCore/SparseDenseProduct.h:26-72
//   ConjugateGradient<...>                   Eigen/src/IterativeLinearSolvers/ConjugateGradient.h:157-225
//   BiCGSTAB<...>                            Eigen/src/IterativeLinearSolvers/BiCGSTAB.h:157-208
//   DiagonalPreconditioner / Identity        Eigen/src/IterativeLinearSolvers/BasicPreconditioners.h:35-108,200-222
//   LeastSquaresConjugateGradient            Eigen/src/IterativeLinearSolvers/LeastSquareConjugateGradient.h
//   MINRES / GMRES                           unsupported/Eigen/src/IterativeSolvers/MINRES.h, GMRES.h
//   IncompleteLUT / IncompleteCholesky       Eigen/src/IterativeLinearSolvers/IncompleteLUT.h, IncompleteCholesky.h
// on caller-owned CSR arrays bound zero-copy through Map<const SparseMatrix> (SparseMap.h:270).
//
// Used for (1) pinning the C restatement in oracle/oracle.c, (2) generating tests/golden/*,
// (3) the `--impl reference` arm and cpu_baseline of bench.py (kind "reference").
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.
#include <Eigen/Sparse>
#include <Eigen/IterativeLinearSolvers>
#include <iostream>  // the unsupported module header uses std::cerr without including it
#include <unsupported/Eigen/IterativeSolvers>
#include <chrono>
#include <cstdint>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace Eigen;
#define COMMA ,

namespace {

template <typename T>
using Csr = SparseMatrix<T, RowMajor, int>;
template <typename T>
using CsrMap = Map<const Csr<T>>;
template <typename T>
using Vec = Matrix<T, Dynamic, 1>;

inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

inline void set_threads(int threads) {
  if (threads > 0) Eigen::setNbThreads(threads);
}

template <typename Solver, typename T>
int run_solver(Solver& solver, const CsrMap<T>& A, int64_t n, const T* b, T* x, int use_guess, double tol,
               int64_t max_iters, int64_t* iters, double* error, int* info, double* seconds_setup,
               double* seconds_solve) {
  Map<const Vec<T>> bm(b, n);
  Map<Vec<T>> xm(x, n);
  double t0 = now_s();
  solver.compute(A);
  double t1 = now_s();
  if (tol >= 0) solver.setTolerance(static_cast<T>(tol));
  if (max_iters >= 0) solver.setMaxIterations(max_iters);
  Vec<T> sol;
  double t2 = now_s();
  if (use_guess) {
    Vec<T> guess = xm;
    sol = solver.solveWithGuess(bm, guess);
  } else {
    sol = solver.solve(bm);
  }
  double t3 = now_s();
  xm = sol;
  if (iters) *iters = solver.iterations();
  if (error) *error = static_cast<double>(solver.error());
  if (info) *info = static_cast<int>(solver.info());
  if (seconds_setup) *seconds_setup = t1 - t0;
  if (seconds_solve) *seconds_solve = t3 - t2;
  return 0;
}

template <typename T>
int cg_dispatch(int64_t n, int64_t nnz, const int* rowptr, const int* colidx, const T* vals, const T* b, T* x,
                int use_guess, double tol, int64_t max_iters, int uplo, int precond, int threads, int64_t* iters,
                double* error, int* info, double* t_setup, double* t_solve) {
  set_threads(threads);
  CsrMap<T> A(n, n, nnz, rowptr, colidx, vals);
#define EIGREF_CG(UPLO, PRE)                                                                          \
  {                                                                                                   \
    ConjugateGradient<Csr<T>, UPLO, PRE> s;                                                           \
    return run_solver(s, A, n, b, x, use_guess, tol, max_iters, iters, error, info, t_setup, t_solve); \
  }
  if (precond == 1) {
    if (uplo == (Lower | Upper)) EIGREF_CG(Lower | Upper, DiagonalPreconditioner<T>)
    if (uplo == Lower) EIGREF_CG(Lower, DiagonalPreconditioner<T>)
    if (uplo == Upper) EIGREF_CG(Upper, DiagonalPreconditioner<T>)
  } else if (precond == 0) {
    if (uplo == (Lower | Upper)) EIGREF_CG(Lower | Upper, IdentityPreconditioner)
    if (uplo == Lower) EIGREF_CG(Lower, IdentityPreconditioner)
    if (uplo == Upper) EIGREF_CG(Upper, IdentityPreconditioner)
  }
#undef EIGREF_CG
  return -1;
}

template <typename T>
int bicgstab_dispatch(int64_t n, int64_t nnz, const int* rowptr, const int* colidx, const T* vals, const T* b,
                      T* x, int use_guess, double tol, int64_t max_iters, int precond, int threads,
                      int64_t* iters, double* error, int* info, double* t_setup, double* t_solve) {
  set_threads(threads);
  CsrMap<T> A(n, n, nnz, rowptr, colidx, vals);
  if (precond == 1) {
    BiCGSTAB<Csr<T>, DiagonalPreconditioner<T>> s;
    return run_solver(s, A, n, b, x, use_guess, tol, max_iters, iters, error, info, t_setup, t_solve);
  } else if (precond == 0) {
    BiCGSTAB<Csr<T>, IdentityPreconditioner> s;
    return run_solver(s, A, n, b, x, use_guess, tol, max_iters, iters, error, info, t_setup, t_solve);
  }
  return -1;
}

template <typename T>
int spmv_impl(int64_t rows, int64_t cols, int64_t nnz, const int* rowptr, const int* colidx, const T* vals,
              const T* x, T* y, int threads, int reps, double* best_seconds) {
  set_threads(threads);
  CsrMap<T> A(rows, cols, nnz, rowptr, colidx, vals);
  Map<const Vec<T>> xm(x, cols);
  Map<Vec<T>> ym(y, rows);
  double best = 1e300;
  for (int r = 0; r < (reps > 0 ? reps : 1); ++r) {
    double t0 = now_s();
    ym.noalias() = A * xm;  // the exact statement of ConjugateGradient.h:71
    double t1 = now_s();
    if (t1 - t0 < best) best = t1 - t0;
  }
  if (best_seconds) *best_seconds = best;
  return 0;
}


// ---- incomplete factorizations as preconditioners ----------------------------------------------
// IncompleteLUT keeps its factor and permutations protected (IncompleteLUT.h:183-189): a derived class only re-exports
// them, no code of the reference is replaced.
struct IlutAccess : IncompleteLUT<double, int> {
  const FactorType& lu() const { return m_lu; }
  const PermutationMatrix<Dynamic, Dynamic, int>& P() const { return m_P; }
  const PermutationMatrix<Dynamic, Dynamic, int>& Pinv() const { return m_Pinv; }
};

template <typename Pre>
void ilut_params(Pre& pre, double droptol, int fillfactor) {
  if (droptol >= 0) pre.setDroptol(droptol);
  if (fillfactor > 0) pre.setFillfactor(fillfactor);
}

template <typename IC>
int64_t ichol_export(const CsrMap<double>& A, double shift, int* colptr, int* rowidx, double* vals, int64_t cap,
                     double* scale, int* perm, int* perm_size, int* info) {
  IC ic;
  if (shift >= 0) ic.setInitialShift(shift);
  ic.compute(A);
  *info = int(ic.info());
  const auto& L = ic.matrixL();
  const int64_t n = L.cols(), nz = L.nonZeros();
  for (int64_t j = 0; j <= n; ++j) colptr[j] = L.outerIndexPtr()[j];
  for (int64_t k = 0; k < std::min<int64_t>(cap, nz); ++k) { rowidx[k] = L.innerIndexPtr()[k]; vals[k] = L.valuePtr()[k]; }
  for (int64_t j = 0; j < ic.scalingS().size(); ++j) scale[j] = ic.scalingS()[j];
  *perm_size = int(ic.permutationP().size());
  for (int64_t j = 0; j < ic.permutationP().size(); ++j) perm[j] = ic.permutationP().indices()[j];
  return nz;
}

template <typename IC>
int ichol_solve(const CsrMap<double>& A, double shift, int64_t n, const double* r, double* z) {
  IC ic;
  if (shift >= 0) ic.setInitialShift(shift);
  ic.compute(A);
  Map<const Vec<double>> rm(r, n);
  Map<Vec<double>> zm(z, n);
  Vec<double> out = ic.solve(rm);
  zm = out;
  return int(ic.info());
}

}  // namespace

extern "C" {

int eigref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

const char* eigref_build_info(void) {
  return "Eigen " EIGEN_MAKESTRING(EIGEN_WORLD_VERSION) "." EIGEN_MAKESTRING(EIGEN_MAJOR_VERSION) "." EIGEN_MAKESTRING(
      EIGEN_MINOR_VERSION) " g++ " __VERSION__
#ifdef __AVX512F__
                           " avx512"
#elif defined(__AVX2__)
                           " avx2"
#else
                           " sse2"
#endif
#ifdef __FMA__
                           " fma"
#endif
#ifdef _OPENMP
                           " openmp"
#endif
      ;
}

int eigref_spmv_f64(int64_t rows, int64_t cols, int64_t nnz, const int* rowptr, const int* colidx,
                    const double* vals, const double* x, double* y, int threads, int reps, double* best_seconds) {
  return spmv_impl<double>(rows, cols, nnz, rowptr, colidx, vals, x, y, threads, reps, best_seconds);
}
int eigref_spmv_f32(int64_t rows, int64_t cols, int64_t nnz, const int* rowptr, const int* colidx,
                    const float* vals, const float* x, float* y, int threads, int reps, double* best_seconds) {
  return spmv_impl<float>(rows, cols, nnz, rowptr, colidx, vals, x, y, threads, reps, best_seconds);
}

// y = selfadjointView<UpLo>(A) * x   (SparseSelfAdjointView.h:279-337); uplo 1=Lower 2=Upper
int eigref_symv_f64(int64_t n, int64_t nnz, const int* rowptr, const int* colidx, const double* vals,
                    const double* x, double* y, int uplo) {
  CsrMap<double> A(n, n, nnz, rowptr, colidx, vals);
  Map<const Vec<double>> xm(x, n);
  Map<Vec<double>> ym(y, n);
  if (uplo == Lower)
    ym.noalias() = A.selfadjointView<Lower>() * xm;
  else if (uplo == Upper)
    ym.noalias() = A.selfadjointView<Upper>() * xm;
  else
    return -1;
  return 0;
}

int eigref_jacobi_f64(int64_t n, int64_t nnz, const int* rowptr, const int* colidx, const double* vals,
                      const double* r, double* z) {
  CsrMap<double> A(n, n, nnz, rowptr, colidx, vals);
  DiagonalPreconditioner<double> pre;
  pre.compute(A);
  Map<const Vec<double>> rm(r, n);
  Map<Vec<double>> zm(z, n);
  zm = pre.solve(rm);
  return 0;
}

int eigref_cg_f64(int64_t n, int64_t nnz, const int* rowptr, const int* colidx, const double* vals,
                  const double* b, double* x, int use_guess, double tol, int64_t max_iters, int uplo, int precond,
                  int threads, int64_t* iters, double* error, int* info, double* t_setup, double* t_solve) {
  return cg_dispatch<double>(n, nnz, rowptr, colidx, vals, b, x, use_guess, tol, max_iters, uplo, precond,
                             threads, iters, error, info, t_setup, t_solve);
}
int eigref_cg_f32(int64_t n, int64_t nnz, const int* rowptr, const int* colidx, const float* vals, const float* b,
                  float* x, int use_guess, double tol, int64_t max_iters, int uplo, int precond, int threads,
                  int64_t* iters, double* error, int* info, double* t_setup, double* t_solve) {
  return cg_dispatch<float>(n, nnz, rowptr, colidx, vals, b, x, use_guess, tol, max_iters, uplo, precond, threads,
                            iters, error, info, t_setup, t_solve);
}
int eigref_bicgstab_f64(int64_t n, int64_t nnz, const int* rowptr, const int* colidx, const double* vals,
                        const double* b, double* x, int use_guess, double tol, int64_t max_iters, int precond,
                        int threads, int64_t* iters, double* error, int* info, double* t_setup,
                        double* t_solve) {
  return bicgstab_dispatch<double>(n, nnz, rowptr, colidx, vals, b, x, use_guess, tol, max_iters, precond,
                                   threads, iters, error, info, t_setup, t_solve);
}
int eigref_bicgstab_f32(int64_t n, int64_t nnz, const int* rowptr, const int* colidx, const float* vals,
                        const float* b, float* x, int use_guess, double tol, int64_t max_iters, int precond,
                        int threads, int64_t* iters, double* error, int* info, double* t_setup, double* t_solve) {
  return bicgstab_dispatch<float>(n, nnz, rowptr, colidx, vals, b, x, use_guess, tol, max_iters, precond, threads,
                                  iters, error, info, t_setup, t_solve);
}

// ---- LSCG (rows x cols matrix), MINRES, GMRES -------------------------------------------------
int eigref_lscg_f64(int64_t rows, int64_t cols, int64_t nnz, const int* rowptr, const int* colidx, const double* vals,
                    const double* b, double* x, int use_guess, double tol, int64_t max_iters, int precond,
                    int64_t* iters, double* error, int* info) {
  CsrMap<double> A(rows, cols, nnz, rowptr, colidx, vals);
  Map<const Vec<double>> bm(b, rows);
  Map<Vec<double>> xm(x, cols);
  Vec<double> sol;
#define EIGREF_LSCG(PRE)                                                        \
  {                                                                             \
    LeastSquaresConjugateGradient<Csr<double>, PRE> s;                          \
    s.compute(A);                                                               \
    if (tol >= 0) s.setTolerance(tol);                                          \
    if (max_iters >= 0) s.setMaxIterations(max_iters);                          \
    if (use_guess) { Vec<double> g = xm; sol = s.solveWithGuess(bm, g); }       \
    else sol = s.solve(bm);                                                     \
    xm = sol;                                                                   \
    *iters = s.iterations(); *error = s.error(); *info = int(s.info());         \
    return 0;                                                                   \
  }
  if (precond == 1) EIGREF_LSCG(LeastSquareDiagonalPreconditioner<double>)
  if (precond == 0) EIGREF_LSCG(IdentityPreconditioner)
#undef EIGREF_LSCG
  return -1;
}

int eigref_minres_f64(int64_t n, int64_t nnz, const int* rowptr, const int* colidx, const double* vals, const double* b,
                      double* x, int use_guess, double tol, int64_t max_iters, int uplo, int precond, int64_t* iters,
                      double* error, int* info) {
  CsrMap<double> A(n, n, nnz, rowptr, colidx, vals);
  double ts, tv;
#define EIGREF_MINRES(UPLO, PRE)                                                                       \
  {                                                                                                    \
    MINRES<Csr<double>, UPLO, PRE> s;                                                                  \
    return run_solver(s, A, n, b, x, use_guess, tol, max_iters, iters, error, info, &ts, &tv);         \
  }
  if (precond == 0) {
    if (uplo == (Lower | Upper)) EIGREF_MINRES(Lower | Upper, IdentityPreconditioner)
    if (uplo == Lower) EIGREF_MINRES(Lower, IdentityPreconditioner)
    if (uplo == Upper) EIGREF_MINRES(Upper, IdentityPreconditioner)
  } else if (precond == 1) {
    if (uplo == (Lower | Upper)) EIGREF_MINRES(Lower | Upper, DiagonalPreconditioner<double>)
    if (uplo == Lower) EIGREF_MINRES(Lower, DiagonalPreconditioner<double>)
    if (uplo == Upper) EIGREF_MINRES(Upper, DiagonalPreconditioner<double>)
  }
#undef EIGREF_MINRES
  return -1;
}

int eigref_gmres_f64(int64_t n, int64_t nnz, const int* rowptr, const int* colidx, const double* vals, const double* b,
                     double* x, int use_guess, double tol, int64_t max_iters, int64_t restart, int precond,
                     int64_t* iters, double* error, int* info) {
  CsrMap<double> A(n, n, nnz, rowptr, colidx, vals);
  double ts, tv;
  if (precond == 1) {
    GMRES<Csr<double>, DiagonalPreconditioner<double>> s;
    if (restart > 0) s.set_restart(restart);
    return run_solver(s, A, n, b, x, use_guess, tol, max_iters, iters, error, info, &ts, &tv);
  } else if (precond == 0) {
    GMRES<Csr<double>, IdentityPreconditioner> s;
    if (restart > 0) s.set_restart(restart);
    return run_solver(s, A, n, b, x, use_guess, tol, max_iters, iters, error, info, &ts, &tv);
  }
  return -1;
}


// ---- SURVEY 8f rank 4 ---------------------------------------------------------------------------------------------
// IncompleteLUT<double>::compute on the row-major matrix: the factor m_lu (row-major, rows NOT sorted: L part, diagonal,
// U part in QuickSplit order) and the AMD permutation m_P.  Returns nnz(m_lu).
int64_t eigref_ilut_f64(int64_t n, int64_t nnz, const int* rowptr, const int* colidx, const double* vals, double droptol,
                        int fillfactor, int* lu_rowptr, int* lu_colidx, double* lu_vals, int64_t cap, int* P, int* Pinv,
                        int* info) {
  CsrMap<double> A(n, n, nnz, rowptr, colidx, vals);
  IlutAccess pre;
  ilut_params(pre, droptol, fillfactor);
  pre.compute(A);
  *info = int(pre.info());
  const auto& lu = pre.lu();
  const int64_t nz = lu.nonZeros();
  for (int64_t i = 0; i <= n; ++i) lu_rowptr[i] = lu.outerIndexPtr()[i];
  for (int64_t k = 0; k < std::min<int64_t>(cap, nz); ++k) { lu_colidx[k] = lu.innerIndexPtr()[k]; lu_vals[k] = lu.valuePtr()[k]; }
  for (int64_t i = 0; i < n; ++i) { P[i] = pre.P().indices()[i]; Pinv[i] = pre.Pinv().indices()[i]; }
  return nz;
}
// z = IncompleteLUT(A).solve(r)   (IncompleteLUT.h:171-176)
int eigref_ilut_solve_f64(int64_t n, int64_t nnz, const int* rowptr, const int* colidx, const double* vals, double droptol,
                          int fillfactor, const double* r, double* z) {
  CsrMap<double> A(n, n, nnz, rowptr, colidx, vals);
  IncompleteLUT<double, int> pre;
  ilut_params(pre, droptol, fillfactor);
  pre.compute(A);
  Map<const Vec<double>> rm(r, n);
  Map<Vec<double>> zm(z, n);
  Vec<double> out = pre.solve(rm);
  zm = out;
  return int(pre.info());
}
// IncompleteCholesky<double, UpLo, Ordering>::compute: m_L (column-major lower), m_scale, m_perm.  ordering 0 natural, 1 AMD.
int64_t eigref_ichol_f64(int64_t n, int64_t nnz, const int* rowptr, const int* colidx, const double* vals, int uplo,
                         int ordering, double shift, int* colptr, int* rowidx, double* lvals, int64_t cap, double* scale,
                         int* perm, int* perm_size, int* info) {
  CsrMap<double> A(n, n, nnz, rowptr, colidx, vals);
  if (uplo == Lower && ordering == 0) return ichol_export<IncompleteCholesky<double, Lower, NaturalOrdering<int>>>(A, shift, colptr, rowidx, lvals, cap, scale, perm, perm_size, info);
  if (uplo == Lower && ordering == 1) return ichol_export<IncompleteCholesky<double, Lower, AMDOrdering<int>>>(A, shift, colptr, rowidx, lvals, cap, scale, perm, perm_size, info);
  if (uplo == Upper && ordering == 0) return ichol_export<IncompleteCholesky<double, Upper, NaturalOrdering<int>>>(A, shift, colptr, rowidx, lvals, cap, scale, perm, perm_size, info);
  if (uplo == Upper && ordering == 1) return ichol_export<IncompleteCholesky<double, Upper, AMDOrdering<int>>>(A, shift, colptr, rowidx, lvals, cap, scale, perm, perm_size, info);
  return -1;
}
int eigref_ichol_solve_f64(int64_t n, int64_t nnz, const int* rowptr, const int* colidx, const double* vals, int uplo,
                           int ordering, double shift, const double* r, double* z) {
  CsrMap<double> A(n, n, nnz, rowptr, colidx, vals);
  if (uplo == Lower && ordering == 0) return ichol_solve<IncompleteCholesky<double, Lower, NaturalOrdering<int>>>(A, shift, n, r, z);
  if (uplo == Lower && ordering == 1) return ichol_solve<IncompleteCholesky<double, Lower, AMDOrdering<int>>>(A, shift, n, r, z);
  if (uplo == Upper && ordering == 0) return ichol_solve<IncompleteCholesky<double, Upper, NaturalOrdering<int>>>(A, shift, n, r, z);
  if (uplo == Upper && ordering == 1) return ichol_solve<IncompleteCholesky<double, Upper, AMDOrdering<int>>>(A, shift, n, r, z);
  return -1;
}
// Solvers with these preconditioners.  which: 0 ConjugateGradient<_, uplo, IncompleteCholesky<double, uplo or Lower, ordering>>,
// 1 BiCGSTAB<_, IncompleteLUT>, 2 GMRES<_, IncompleteLUT>, 3 MINRES<_, uplo, IncompleteCholesky<...>> is not a reference
// combination and is not offered.
int eigref_precond_solver_f64(int which, int64_t n, int64_t nnz, const int* rowptr, const int* colidx, const double* vals,
                              const double* b, double* x, int use_guess, double tol, int64_t max_iters, int uplo,
                              int ordering, double droptol, int fillfactor, int64_t restart, int64_t* iters,
                              double* error, int* info) {
  CsrMap<double> A(n, n, nnz, rowptr, colidx, vals);
  double ts, tv;
#define EIGREF_RUN(SOLVER, SETUP)                                                                     \
  {                                                                                                    \
    SOLVER s;                                                                                          \
    SETUP;                                                                                             \
    return run_solver(s, A, n, b, x, use_guess, tol, max_iters, iters, error, info, &ts, &tv);         \
  }
  typedef Csr<double> M;
  if (which == 0) {
    if (uplo == Lower && ordering == 0) EIGREF_RUN(ConjugateGradient<M COMMA Lower COMMA IncompleteCholesky<double COMMA Lower COMMA NaturalOrdering<int>>>, (void)0)
    if (uplo == Lower && ordering == 1) EIGREF_RUN(ConjugateGradient<M COMMA Lower COMMA IncompleteCholesky<double COMMA Lower COMMA AMDOrdering<int>>>, (void)0)
    if (uplo == Upper && ordering == 0) EIGREF_RUN(ConjugateGradient<M COMMA Upper COMMA IncompleteCholesky<double COMMA Upper COMMA NaturalOrdering<int>>>, (void)0)
    if (uplo == Upper && ordering == 1) EIGREF_RUN(ConjugateGradient<M COMMA Upper COMMA IncompleteCholesky<double COMMA Upper COMMA AMDOrdering<int>>>, (void)0)
    if (uplo == (Lower | Upper) && ordering == 0) EIGREF_RUN(ConjugateGradient<M COMMA Lower | Upper COMMA IncompleteCholesky<double COMMA Lower COMMA NaturalOrdering<int>>>, (void)0)
    if (uplo == (Lower | Upper) && ordering == 1) EIGREF_RUN(ConjugateGradient<M COMMA Lower | Upper COMMA IncompleteCholesky<double COMMA Lower COMMA AMDOrdering<int>>>, (void)0)
  } else if (which == 1) {
    EIGREF_RUN(BiCGSTAB<M COMMA IncompleteLUT<double>>, ilut_params(s.preconditioner(), droptol, fillfactor))
  } else if (which == 2) {
    EIGREF_RUN(GMRES<M COMMA IncompleteLUT<double>>, (ilut_params(s.preconditioner(), droptol, fillfactor), restart > 0 ? s.set_restart(restart) : (void)0))
  }
#undef EIGREF_RUN
  return -1;
}

}  // extern "C"

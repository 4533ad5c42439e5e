// oracle/ordering_check.cpp -- TEST INFRASTRUCTURE (CPU; built and run in the build container by tests/test_ordering_cpp.py).
//
// b200::MulticolorOrdering (include/b200/Ordering.h) inside the reference's own CPU classes: Eigen::IncompleteCholesky
// with that ordering factorizes, Eigen::ConjugateGradient with it converges to the solution of a direct solve, and the
// factor handed to the GPU-free analysis (b200s_factors_from_ichol_f64) has two dependency levels per triangular solve on
// the 7-point stencil (the natural ordering: 3n - 2).  No device is touched.
#include <Eigen/IterativeLinearSolvers>
#include <Eigen/SparseCholesky>
#include <Eigen/SparseCore>

#include <cstdio>
#include <vector>

#include <b200/IterativeSolvers.h>
#include <b200/Ordering.h>

typedef Eigen::SparseMatrix<double> SpMat;

static SpMat poisson3d(int n) {
  std::vector<Eigen::Triplet<double> > t;
  for (int k = 0; k < n; ++k)
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        const int r = i + n * (j + n * k);
        t.push_back(Eigen::Triplet<double>(r, r, 6.0));
        if (i > 0) t.push_back(Eigen::Triplet<double>(r, r - 1, -1.0));
        if (i < n - 1) t.push_back(Eigen::Triplet<double>(r, r + 1, -1.0));
        if (j > 0) t.push_back(Eigen::Triplet<double>(r, r - n, -1.0));
        if (j < n - 1) t.push_back(Eigen::Triplet<double>(r, r + n, -1.0));
        if (k > 0) t.push_back(Eigen::Triplet<double>(r, r - n * n, -1.0));
        if (k < n - 1) t.push_back(Eigen::Triplet<double>(r, r + n * n, -1.0));
      }
  SpMat A(n * n * n, n * n * n);
  A.setFromTriplets(t.begin(), t.end());
  return A;
}

template <typename IC>
static int levels_of(const IC& ic, int32_t out[2]) {
  b200s_factors* f = b200::detail::factors_of<IC>::make(ic);
  if (!f) return 1;
  for (int w = 0; w < 2; ++w) b200s_factors_stage_sizes(f, w, 0, &out[w], 0, 0, 0);
  b200s_factors_destroy(f);
  return 0;
}

int main() {
  const int n = 12;
  SpMat A = poisson3d(n);
  Eigen::VectorXd xt = Eigen::VectorXd::Random(A.rows()), b = A * xt;
  typedef Eigen::IncompleteCholesky<double, Eigen::Lower, b200::MulticolorOrdering<int> > IcMc;
  typedef Eigen::IncompleteCholesky<double, Eigen::Lower, Eigen::NaturalOrdering<int> > IcNat;
  Eigen::ConjugateGradient<SpMat, Eigen::Lower, IcMc> cg(A);
  cg.setTolerance(1e-12);
  Eigen::VectorXd x = cg.solve(b);
  if (cg.info() != Eigen::Success || cg.preconditioner().info() != Eigen::Success) return std::printf("FAIL: CG+IC(multicolour) info\n"), 1;
  if ((x - xt).norm() > 1e-9 * xt.norm()) return std::printf("FAIL: solution %g\n", (x - xt).norm() / xt.norm()), 1;
  const Eigen::Index ps = cg.preconditioner().permutationP().size();
  if (ps != A.rows()) return std::printf("FAIL: permutation size %ld\n", long(ps)), 1;
  int32_t lev[2], nat[2];
  if (levels_of(cg.preconditioner(), lev)) return std::printf("FAIL: factors_of: %s\n", b200s_last_error(0)), 1;
  IcNat icn(A);
  if (levels_of(icn, nat)) return std::printf("FAIL: factors_of(natural): %s\n", b200s_last_error(0)), 1;
  std::printf("multicolour: %d iterations, levels %d %d; natural ordering levels %d %d\n", int(cg.iterations()), lev[0],
              lev[1], nat[0], nat[1]);
  if (lev[0] != 2 || lev[1] != 2 || nat[0] != 3 * n - 2 || nat[1] != 3 * n - 2) return std::printf("FAIL: levels\n"), 1;
  // the general-matrix overload (what an LU-type preconditioner would call)
  b200::MulticolorOrdering<int> ord;
  b200::MulticolorOrdering<int>::PermutationType p;
  ord(A, p);
  if (ord.colours() != 2 || p.size() != A.rows()) return std::printf("FAIL: general overload\n"), 1;
  std::printf("ordering_check ok\n");
  return 0;
}

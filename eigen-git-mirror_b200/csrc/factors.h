// factors.h -- host-side (GPU-free) half of the incomplete-factorization preconditioners.
//
// The reference applies IncompleteLUT / IncompleteCholesky in every solver iteration as two sparse triangular solves
// between permutations (IncompleteLUT.h:171-176, IncompleteCholesky.h:149-157, TriangularSolver.h:26-134).  That
// application is the per-iteration hot part and runs on the GPU (kernels_tri.cuh): each solve is split into dependency
// LEVELS, the rows of one level are independent and are handled by one thread each, which sums its row in the very
// order the reference's substitution loop does (same roundings per entry, one IEEE division) -- so z = M^-1 r has the
// bits of the reference as g++ -O3 compiles it for an FMA-capable x86-64.
//
// The factorization itself is sequential by construction in the reference as well (row ii of ILUT needs rows < ii);
// it is setup work, done once per matrix on the host by the restatements in factors.cpp, or by the caller (the C++
// binding hands over the factor an Eigen::IncompleteLUT / IncompleteCholesky object already holds).
// No CUDA call is made in this translation unit (tests/test_factors.py runs on a CPU-only box).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/b200sparse.h"

namespace b200s {

// One triangular solve T x = y, done in place.  Row i: t = x[i]; t -= vals[k] * x[colidx[k]] for the row's entries k
// in storage order; x[i] = unit ? t : t / diag[i].  `fused` says how each step rounds: the reference's column sweep
// (`x[i] -= tmp * value`, TriangularSolver.h:129) is contracted into one FMA by g++ -O3 when FMA is available, its
// row-wise loops (`tmp -= value * x[col]`, :49 and :91) are not (probed on the unmodified reference, tests/test_factors.py).
struct TriStage {
  bool fused = false;
  std::vector<int32_t> rowptr, colidx;  // off-diagonal entries per row, in the reference's summation order
  std::vector<double> vals;
  std::vector<double> diag;             // empty = unit diagonal
  // dependency levels: rows level_rows[level_ptr[l] .. level_ptr[l+1]) only read x of rows in levels < l
  std::vector<int32_t> level_ptr, level_rows;
  // launch plan: consecutive levels that each fit one CTA are fused into one single-CTA launch
  struct Launch { int32_t level_begin, level_end, rows; };
  std::vector<Launch> launches;
};

struct Factors {
  int kind = 0;      // B200S_FACTORS_ILUT / B200S_FACTORS_ICHOL
  int64_t n = 0;
  int info = 0;      // Eigen's ComputationInfo of the factorization (0 Success, 1 NumericalIssue)
  // M^-1 r:  x[k] = pre_scale[k] * r[pre_gather[k]];  first;  second;  z[k] = post_scale[k] * x[post_gather[k]]
  // (empty gather = identity, empty scale = 1)
  std::vector<int32_t> pre_gather, post_gather;
  std::vector<double> pre_scale, post_scale;
  TriStage first, second;
  // the factor as the reference stores it: ILUT m_lu (row-major, n+1 / nnz), ICHOL m_L (column-major lower)
  std::vector<int32_t> outer, inner;
  std::vector<double> vals;
  std::vector<double> scale;   // ICHOL m_scale
  std::vector<int32_t> perm;   // ILUT m_P.indices() (always n entries), ICHOL m_perm.indices() (empty = natural)
};

constexpr int kTriBlock = 256;         // threads per CTA of a wide level
constexpr int kTriFusedBlock = 1024;   // a run of levels with <= this many rows each is one single-CTA launch

// IncompleteLUT<double>::factorize (IncompleteLUT.h:238-446) on the CSR matrix `a`, with the fill-reducing
// permutation given by the caller: perm = m_P.indices() (NULL = identity; the reference computes it with AMD,
// IncompleteLUT.h:221-236, which is an ordering heuristic outside this path).  droptol < 0 -> 1e-12, fillfactor <= 0 -> 10.
int ilut_factorize(int64_t n, const int32_t* rowptr, const int32_t* colidx, const double* vals, double droptol,
                   int fillfactor, const int32_t* perm, Factors& f, std::string& err);
// IncompleteCholesky<double, UpLo, Ordering>::factorize (IncompleteCholesky.h:200-367); perm = m_perm.indices()
// (NULL = NaturalOrdering).  shift < 0 -> 1e-3.
int ichol_factorize(int64_t n, const int32_t* rowptr, const int32_t* colidx, const double* vals, int uplo, double shift,
                    const int32_t* perm, Factors& f, std::string& err);
// Factors computed elsewhere (an Eigen::IncompleteLUT's m_lu / m_P, an Eigen::IncompleteCholesky's m_L / m_scale / m_perm).
int factors_from_ilut(int64_t n, const int32_t* lu_rowptr, const int32_t* lu_colidx, const double* lu_vals,
                      const int32_t* perm, Factors& f, std::string& err);
int factors_from_ichol(int64_t n, const int32_t* colptr, const int32_t* rowidx, const double* lvals, const double* scale,
                       const int32_t* perm, Factors& f, std::string& err);

// A fill-reducing ordering is not what a GPU wants from the permutation: the triangular solves run level by level, so
// the ordering that matters here is one with FEW, WIDE dependency levels.  Greedy multi-colouring of the symmetrised
// pattern (vertices in natural order, smallest free colour), then vertices sorted by colour: within a colour no two
// rows are coupled, so a zero-fill factor has one level per colour (red-black for the 5/7-point stencils: 2 levels
// instead of ~3n).  perm is in the reference's convention (row i of A becomes row perm[i]); returns the colour count.
int multicolor_ordering(int64_t n, const int32_t* rowptr, const int32_t* colidx, int32_t* perm, std::string& err);

}  // namespace b200s

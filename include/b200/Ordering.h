// b200/Ordering.h -- an ordering functor in the reference's Ordering concept (Eigen/src/OrderingMethods/Ordering.h:48-107)
// that orders for the GPU rather than for fill.
//
//   b200::MulticolorOrdering<StorageIndex>   usable wherever AMDOrdering / NaturalOrdering are, e.g.
//       b200::ConjugateGradient<SpMat, Eigen::Lower,
//                               Eigen::IncompleteCholesky<double, Eigen::Lower, b200::MulticolorOrdering<int> > > cg(A);
//
// The incomplete-factorization preconditioners are applied on the device level by level (b200sparse.h): the ordering that
// matters there has FEW, WIDE dependency levels.  Greedy multi-colouring of the pattern of A + A^T, rows sorted by colour
// (b200s_ordering_multicolor, host): rows of one colour are not coupled, so a zero-fill factor has one level per colour --
// red-black on the 5/7-point stencils, 2 levels instead of ~3n -- at the price of some extra iterations.
// As with the reference's functors the result is the permutation whose inverse the factorization stores
// (IncompleteCholesky.h:95-105: m_perm = pinv.inverse()).
#ifndef B200_ORDERING_H
#define B200_ORDERING_H

#include <Eigen/SparseCore>

#include <vector>

#include "../b200sparse.h"

namespace b200 {

template <typename StorageIndex>
class MulticolorOrdering {
 public:
  typedef Eigen::PermutationMatrix<Eigen::Dynamic, Eigen::Dynamic, StorageIndex> PermutationType;

  /** The permutation from the pattern of a general sparse matrix (A + A^T is what gets coloured). */
  template <typename MatrixType>
  void operator()(const MatrixType& mat, PermutationType& perm) {
    Eigen::SparseMatrix<typename MatrixType::Scalar, Eigen::RowMajor, StorageIndex> C = mat;
    run(C, perm);
  }
  /** The permutation from a self-adjoint view (what IncompleteCholesky::analyzePattern passes, IncompleteCholesky.h:98). */
  template <typename SrcType, unsigned int SrcUpLo>
  void operator()(const Eigen::SparseSelfAdjointView<SrcType, SrcUpLo>& mat, PermutationType& perm) {
    Eigen::SparseMatrix<typename SrcType::Scalar, Eigen::RowMajor, StorageIndex> C;
    C = mat;
    run(C, perm);
  }
  int colours() const { return m_colours; }

 private:
  template <typename Csr>
  void run(Csr& C, PermutationType& perm) {
    C.makeCompressed();
    const Eigen::Index n = C.rows();
    std::vector<int32_t> rp(C.outerIndexPtr(), C.outerIndexPtr() + n + 1), ci(C.innerIndexPtr(), C.innerIndexPtr() + C.nonZeros());
    std::vector<int32_t> to(n > 0 ? n : 1);
    m_colours = b200s_ordering_multicolor(n, rp.data(), ci.data(), to.data());  // row i goes to position to[i]
    perm.resize(n);
    if (m_colours < 0) {  // cannot happen for a well-formed matrix; fall back to the identity rather than to garbage
      for (Eigen::Index i = 0; i < n; ++i) perm.indices()[i] = StorageIndex(i);
      return;
    }
    // the factorization stores the INVERSE of what the functor returns (m_perm = pinv.inverse()), and m_perm must send
    // row i to to[i]: return the permutation with indices()[to[i]] = i
    for (Eigen::Index i = 0; i < n; ++i) perm.indices()[to[i]] = StorageIndex(i);
  }
  int m_colours = 0;
};

}  // namespace b200

#endif  // B200_ORDERING_H

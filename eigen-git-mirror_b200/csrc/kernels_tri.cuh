// kernels_tri.cuh -- applying an incomplete factorization on the GPU:
//   IncompleteLUT::_solve_impl       x = Pinv b;  L\x (unit lower);  U\x;  z = P x         IncompleteLUT.h:171-176
//   IncompleteCholesky::_solve_impl  x = perm b;  x = S x;  L\x;  L^T\x;  x = S x;  z = perm^-1 x   IncompleteCholesky.h:149-157
// The substitution loops of the reference (TriangularSolver.h:26-134) are sequential over rows; here the rows are grouped
// into dependency levels on the host (csrc/factors.cpp): the rows of one level only read entries of x finished by earlier
// levels, so one thread per row sums its row exactly as the reference's inner loop does -- entry by entry in storage
// order, each step rounded as the reference's compiled loop rounds it (TriArgs::fused), one IEEE division by the
// diagonal -- and the result does not depend on how rows are spread over threads.  HBM traffic per apply: both factors once (12 B per stored entry) plus the gathers
// of x; a bandwidth/latency-bound kernel like the SpMV, with far less parallelism per launch (one level).
#pragma once
#include <cstdint>

namespace b200s {

struct TriArgs {
  const int32_t* rowptr;
  const int32_t* colidx;
  const double* vals;
  const double* diag;        // nullptr = unit diagonal
  const int32_t* level_ptr;
  const int32_t* level_rows;
  double* x;                 // solved in place; read AND written inside one launch, so never through the read-only path
  int fused;                 // 1: t = fma(-v, x, t);  0: product and subtraction rounded separately (csrc/factors.h)
};

__device__ __forceinline__ void tri_solve_row(const TriArgs& a, int row) {
  double t = a.x[row];
  const int e = a.rowptr[row + 1];
  if (a.fused) {
    for (int k = a.rowptr[row]; k < e; ++k) t = fma(-a.vals[k], a.x[a.colidx[k]], t);
  } else {
    for (int k = a.rowptr[row]; k < e; ++k) t = __dsub_rn(t, __dmul_rn(a.vals[k], a.x[a.colidx[k]]));  // never contracted
  }
  a.x[row] = a.diag ? t / a.diag[row] : t;
}

// Levels [level_begin, level_end).  A single wide level runs on a whole grid; a run of narrow levels runs on ONE CTA
// (the host launches it with gridDim.x == 1), separated by block barriers: __syncthreads orders the global writes of
// level l before the reads of level l+1 for the threads of the block.
__global__ void __launch_bounds__(1024) tri_levels_kernel(const TriArgs a, int level_begin, int level_end) {
  const int stride = gridDim.x * blockDim.x;
  for (int l = level_begin; l < level_end; ++l) {
    const int end = a.level_ptr[l + 1];
    for (int k = a.level_ptr[l] + blockIdx.x * blockDim.x + threadIdx.x; k < end; k += stride)
      tri_solve_row(a, a.level_rows[k]);
    if (level_end - level_begin > 1) __syncthreads();
  }
}

// out[k] = scale[k] * in[gather[k]]   (gather == nullptr: identity, scale == nullptr: 1); `out` never aliases `in`
__global__ void __launch_bounds__(256) tri_permute_scale_kernel(long long n, const double* __restrict__ in,
                                                                const int32_t* __restrict__ gather,
                                                                const double* __restrict__ scale,
                                                                double* __restrict__ out) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += stride) {
    const double v = in[gather ? gather[k] : k];
    out[k] = scale ? scale[k] * v : v;
  }
}

}  // namespace b200s

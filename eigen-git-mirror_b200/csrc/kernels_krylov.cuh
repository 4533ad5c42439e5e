// kernels_krylov.cuh -- vector primitives for the solvers that reuse the hot path's SpMV / axpy / dot
// (SURVEY 8f rank 3): LeastSquaresConjugateGradient, MINRES, GMRES.  Their loops run on the host side of the library
// (csrc/krylov.inc) exactly as written in the reference, one kernel per vector statement and one deterministic
// reduction per dot product; all vectors stay in device memory.
#pragma once
#include "kernels.cuh"

namespace b200s {

struct DotArgs {
  long long n;
  const double* x0;
  const double* y0;
  const double* x1;  // second pair (nullptr: one dot)
  const double* y1;
  double* out;       // [2] device
  double* partials;
  unsigned int* counter;
};

// out[0] = x0.y0, out[1] = x1.y1 -- same deterministic shape as every other reduction of the library
__global__ void __launch_bounds__(kVecThreads) krylov_dot_kernel(const DotArgs a) {
  __shared__ double scratch[32 * 2];
  __shared__ int s_last_k;
  double v[2] = {0.0, 0.0};
  const bool two = a.x1 != nullptr;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    v[0] = fma_rn(a.x0[i], a.y0[i], v[0]);
    if (two) v[1] = fma_rn(a.x1[i], a.y1[i], v[1]);
  }
  block_reduce<2, kVecThreads>(v, scratch);
  if (threadIdx.x == 0) {
    a.partials[blockIdx.x * 4 + 0] = v[0];
    a.partials[blockIdx.x * 4 + 1] = v[1];
    __threadfence();
    s_last_k = (atomicAdd(a.counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last_k) return;
  __threadfence();
  double t[2] = {0.0, 0.0};
  for (unsigned b = threadIdx.x; b < gridDim.x; b += kVecThreads) {
    t[0] += __ldcg(a.partials + b * 4 + 0);
    t[1] += __ldcg(a.partials + b * 4 + 1);
  }
  block_reduce<2, kVecThreads>(t, scratch);
  if (threadIdx.x == 0) {
    *a.counter = 0;
    a.out[0] = t[0];
    a.out[1] = t[1];
  }
}

// z = a*x + b*y   (x or y may be nullptr = absent term; z may alias x or y)
__global__ void __launch_bounds__(kVecThreads) krylov_axpby_kernel(long long n, double a, const double* x, double b,
                                                                   const double* y, double* z) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    double r = 0.0;
    if (x && y) r = fma_rn(a, x[i], b * y[i]);
    else if (x) r = a * x[i];
    else if (y) r = b * y[i];
    z[i] = r;
  }
}

// z = d .* r  (preconditioner apply, BasicPreconditioners.h:91)
__global__ void __launch_bounds__(kVecThreads) krylov_mul_kernel(long long n, const double* d, const double* r, double* z) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) z[i] = d[i] * r[i];
}

// z = (w - r2*p_old - r3*p_oold) / r1   (MINRES.h:123)
__global__ void __launch_bounds__(kVecThreads) krylov_minres_p_kernel(long long n, const double* w, double r2,
                                                                      const double* p_old, double r3,
                                                                      const double* p_oold, double r1, double* z) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    z[i] = (w[i] - r2 * p_old[i] - r3 * p_oold[i]) / r1;
}

// LeastSquareDiagonalPreconditioner (BasicPreconditioners.h:152-177) from the rows of A^T: invdiag[j] =
// 1 / sum_i |A_ij|^2 when that sum is positive, else `otherwise` (0 for a row-major A, 1 for a column-major A -- the
// reference's two branches differ)
__global__ void lscg_invdiag_kernel(int rows_t, const int32_t* __restrict__ rowptr_t, const double* __restrict__ vals_t,
                                    double otherwise, double* __restrict__ invdiag) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= rows_t) return;
  double s = 0.0;
  for (int k = rowptr_t[j]; k < rowptr_t[j + 1]; ++k) s = fma_rn(vals_t[k], vals_t[k], s);
  invdiag[j] = (s > 0.0) ? 1.0 / s : otherwise;
}

}  // namespace b200s

// kernels_multi.cuh -- K right-hand sides at once: the CG loop of ConjugateGradient.h:26-91 for K columns that share
// ONE stream of the matrix per iteration.
//
// The reference solves a multi-column right-hand side as K sequential solves (IterativeSolverBase.h:375-388), i.e. it
// reads the matrix K times per iteration index; its row-major multi-column product is SparseDenseProduct.h:109-147.
// Here the K columns are stored interleaved (row i holds its K values contiguously, so one gather of a column index
// fetches K operands: 32 bytes for K = 4, exactly the sector a single 8-byte gather occupies anyway), every kernel of
// the single-column pipeline has a K-wide twin, and each column keeps its own scalar recurrence (alpha, beta,
// thresholds, iteration counter, stop flag: one Scalars block per column).  A column that has stopped is frozen while
// the others continue.
//
// Per-column results are BIT-IDENTICAL to the single-column solve: every thread handles the same rows as in the
// single-column kernels (same tiles, same grid, same element-to-thread mapping), performs the same operations in the
// same order for each column, and the block / grid folds have the same shape.
#pragma once
#include "kernels.cuh"

namespace b200s {

constexpr int kMultiMax = 8;                 // columns per batch
constexpr int kMultiPartialStride = 3 * kMultiMax;  // doubles per CTA in the partials buffer of the K-wide kernels

template <int K>
struct MultiArgs {
  long long n;
  double* x;        // [n][K]
  double* r;        // [n][K]
  double* p;        // [n][K]
  double* q;        // [n][K]  A p
  const double* b;  // [n][K]
  const double* invdiag;  // [n]
  Scalars* S;       // [K] one control block per column; S[0].stop_all drives the loop
  double* partials; // [grid][kMultiPartialStride]
  unsigned int* counter;
  unsigned long long cond_handle;
  int set_cond;
  int epilogue;
  int gate;         // kGateLoop: skip when every column has stopped; kGateGuess: run only with an initial guess
};

__device__ __forceinline__ bool multi_gated_out(const Scalars* S, int gate) {
  switch (gate) {
    case kGateLoop: return S[0].stop_all != 0;
    case kGateGuess: return S[0].use_guess == 0;
    default: return false;
  }
}

// Same shape as finish_reduction: block tree -> per-CTA partial -> the last CTA folds all partials in the same thread
// pattern -> per-column scalar epilogue.  v is laid out [column][value]: column j owns v[j*NVC .. j*NVC+NVC).
template <int NVC, int K, int THREADS>
__device__ __forceinline__ void finish_reduction_multi(const MultiArgs<K>& a, double (&v)[NVC * K], double* scratch) {
  constexpr int NV = NVC * K;
  __shared__ int s_last_m;
  block_reduce<NV, THREADS>(v, scratch);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int j = 0; j < NV; ++j) a.partials[blockIdx.x * kMultiPartialStride + j] = v[j];
    __threadfence();
    const unsigned ticket = atomicAdd(a.counter, 1u);
    s_last_m = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last_m) return;
  __threadfence();
  double t[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) t[j] = 0.0;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += THREADS)
#pragma unroll
    for (int j = 0; j < NV; ++j) t[j] += __ldcg(a.partials + b * kMultiPartialStride + j);
  block_reduce<NV, THREADS>(t, scratch);
  if (threadIdx.x == 0) {
    *a.counter = 0;
    int all_stop = 1;
    for (int j = 0; j < K; ++j) {
      RedCtx ctx{};
      ctx.S = a.S + j;
      ctx.epilogue = a.epilogue;
      double r[4] = {0, 0, 0, 0};
      for (int i = 0; i < NVC; ++i) r[i] = t[j * NVC + i];
      const int was_stopped = a.S[j].stop;
      // a column that has stopped keeps its state: its reduced values are stale copies of frozen vectors
      if (a.epilogue == kEpiCgInit || !was_stopped) run_epilogue(ctx, r, nullptr);
      all_stop &= (a.S[j].stop != 0);
    }
    a.S[0].stop_all = all_stop;
    if (a.set_cond) cudaGraphSetConditional(static_cast<cudaGraphConditionalHandle>(a.cond_handle), all_stop ? 0u : 1u);
  }
}

// ------------------------------------------------------------------------------------------------ K-wide SpMV
// y[i][:] = sum_k A[i,k] x[k][:]  (+ per-column dots w.y); same tiles, same lanes per row, same accumulation order
// per column as tile_rows_reduce.
// Row access: the K values of a row are contiguous.  K = 2: one 128-bit access; K >= 4: 256-bit accesses (LDG.E.256 /
// STG.E.256 on sm_100): one instruction per 32-byte sector, so a warp that walks consecutive rows touches every
// 128-byte line once per instruction -- two 128-bit loads per row would double the L1 wavefronts (measured: the
// K = 4 product was L1tex-bound that way).
template <int K>
__device__ __forceinline__ void ldrow(const double* p, long long row, double (&out)[K]) {
  const double* q = p + row * K;
  if (K == 2) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(q));
    out[0] = t.x;
    out[1] = t.y;
  } else {
#pragma unroll
    for (int j = 0; j < K; j += 4)
      asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];"
                   : "=d"(out[j]), "=d"(out[j + 1]), "=d"(out[j + 2]), "=d"(out[j + 3])
                   : "l"(q + j));
  }
}

template <int K>
__device__ __forceinline__ void ldk(const double* p, long long row, double (&out)[K]) {
  const double* q = p + row * K;
  if (K == 2) {
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(out[0]), "=d"(out[1]) : "l"(q) : "memory");
  } else {
#pragma unroll
    for (int j = 0; j < K; j += 4)
      asm volatile("ld.global.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];"
                   : "=d"(out[j]), "=d"(out[j + 1]), "=d"(out[j + 2]), "=d"(out[j + 3])
                   : "l"(q + j)
                   : "memory");
  }
}
template <int K>
__device__ __forceinline__ void stk(double* p, long long row, const double (&v)[K]) {
  double* q = p + row * K;
  if (K == 2) {
    *reinterpret_cast<double2*>(q) = make_double2(v[0], v[1]);
  } else {
#pragma unroll
    for (int j = 0; j < K; j += 4)
      asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(q + j), "d"(v[j]), "d"(v[j + 1]), "d"(v[j + 2]),
                   "d"(v[j + 3])
                   : "memory");
  }
}

template <int K, int LG>
__device__ __forceinline__ void tile_rows_reduce_k(const double* __restrict__ sv, const int32_t* __restrict__ sc,
                                                   const int32_t* __restrict__ srp, int nrows, int row0, int vb0,
                                                   int cb0, const double* x, double* __restrict__ y, const double* w,
                                                   double (&d0)[K]) {
  constexpr int L = 1 << LG;
  const int lane = threadIdx.x & (L - 1);
  const int grp = threadIdx.x >> LG;
  constexpr int NGRP = kSpmvThreads >> LG;
  constexpr int B = (K >= 8) ? 1 : (K == 4) ? 2 : 4;  // entries per lane and trip: B*K = 8 gathered doubles in flight per thread
  for (int base = 0; base < nrows; base += NGRP) {
    const int r = base + grp;
    double sum[K];
#pragma unroll
    for (int q = 0; q < K; ++q) sum[q] = 0.0;
    if (r < nrows) {
      int k = srp[r];
      const int k1 = srp[r + 1];
      for (k += lane; k < k1; k += B * L) {
        int c[B];
        double v[B], xv[B][K];
#pragma unroll
        for (int j = 0; j < B; ++j) {
          const bool in = (k + j * L) < k1;
          c[j] = in ? sc[k + j * L - cb0] : 0;
          v[j] = in ? sv[k + j * L - vb0] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < B; ++j) {
          if ((k + j * L) < k1) {
            ldrow<K>(x, c[j], xv[j]);
          } else {
#pragma unroll
            for (int q = 0; q < K; ++q) xv[j][q] = 0.0;
          }
        }
#pragma unroll
        for (int j = 0; j < B; ++j)
#pragma unroll
          for (int q = 0; q < K; ++q) sum[q] = add_rn(sum[q], mul_rn(v[j], xv[j][q]));
      }
    }
    if (L > 1) {
#pragma unroll
      for (int o = L / 2; o > 0; o >>= 1)
#pragma unroll
        for (int q = 0; q < K; ++q) sum[q] = add_rn(sum[q], __shfl_xor_sync(0xffffffffu, sum[q], o));
    }
    if (r < nrows && lane == 0) {
      double wr[K];
      ldrow<K>(w, row0 + r, wr);
#pragma unroll
      for (int q = 0; q < K; ++q) {
        sum[q] = add_rn(sum[q], 0.0);  // -0 -> +0 as in the single-column product
        d0[q] = fma_rn(wr[q], sum[q], d0[q]);
      }
      stk<K>(y, row0 + r, sum);
    }
  }
}

template <int K>
struct SpmmArgs {
  SpmvArgs<double> sp;  // tiles, pattern, values, stage geometry (x / y / red unused)
  MultiArgs<K> m;
  const double* x;      // [cols][K]
  double* y;            // [rows][K]
  int ndot;             // 0: plain product, 1: per-column dot x.y into the epilogue
};

template <int K>
__global__ void __launch_bounds__(kSpmvThreads) spmm_staged_kernel(const SpmmArgs<K> a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[8];
  __shared__ double red_scratch[32 * K];
  __shared__ double long_scratch[32];
  SpmvCta<double> cx;
  spmv_cta_init(a.sp, cx, smem, full_bar, long_scratch);
  spmv_prefetch(a.sp, cx);
  pdl_wait();
  if (multi_gated_out(a.m.S, a.m.gate)) {  // let the copies land, then leave
    for (int k = 0; k < a.sp.stages; ++k)
      if (static_cast<int>(blockIdx.x + k * gridDim.x) < a.sp.ntiles) mbar_wait(&full_bar[k], 0);
    return;
  }
  double d0[K];
#pragma unroll
  for (int q = 0; q < K; ++q) d0[q] = 0.0;
  const int S = a.sp.stages, G = gridDim.x, tid = threadIdx.x;
  Tile tl_next = (static_cast<int>(blockIdx.x) < a.sp.ntiles) ? a.sp.tiles[blockIdx.x] : Tile{};
  for (int k = 0;; ++k) {
    const int t = blockIdx.x + k * G;
    if (t >= a.sp.ntiles) break;
    const int s = k % S;
    const unsigned parity = (k / S) & 1;
    const Tile tl = tl_next;  // descriptors one round ahead, as in spmv_tiles
    if (t + G < a.sp.ntiles) tl_next = a.sp.tiles[t + G];
    const bool refill = (tid == 0) && (t + S * G < a.sp.ntiles);
    Tile tl_refill = tl;
    if (refill) tl_refill = a.sp.tiles[t + S * G];
    const int nrows = tl.meta & 0xFFFF;
    const int lg = (tl.meta >> 16) & 0xFF;
    mbar_wait(&full_bar[s], parity);
    unsigned char* st = cx.smem + static_cast<size_t>(s) * cx.stage_bytes;
    const double* sv = reinterpret_cast<const double*>(st);
    const int32_t* sc = reinterpret_cast<const int32_t*>(st + cx.v_bytes);
    const int32_t* srp = reinterpret_cast<const int32_t*>(st + cx.v_bytes + cx.c_bytes) + (tl.row0 - (tl.row0 & ~3));
    const int vb0 = tl.nnz0 & ~1, cb0 = tl.nnz0 & ~3;
    switch (lg) {  // the plan admits only row-lane tiles here (no two-phase / long-row tiles)
      case 0: tile_rows_reduce_k<K, 0>(sv, sc, srp, nrows, tl.row0, vb0, cb0, a.x, a.y, a.x, d0); break;
      case 1: tile_rows_reduce_k<K, 1>(sv, sc, srp, nrows, tl.row0, vb0, cb0, a.x, a.y, a.x, d0); break;
      case 2: tile_rows_reduce_k<K, 2>(sv, sc, srp, nrows, tl.row0, vb0, cb0, a.x, a.y, a.x, d0); break;
      case 3: tile_rows_reduce_k<K, 3>(sv, sc, srp, nrows, tl.row0, vb0, cb0, a.x, a.y, a.x, d0); break;
      case 4: tile_rows_reduce_k<K, 4>(sv, sc, srp, nrows, tl.row0, vb0, cb0, a.x, a.y, a.x, d0); break;
      default: tile_rows_reduce_k<K, 5>(sv, sc, srp, nrows, tl.row0, vb0, cb0, a.x, a.y, a.x, d0); break;
    }
    __syncthreads();
    if (refill) spmv_issue_tile<double>(a.sp, cx, tl_refill, s);
  }
  if (a.ndot) finish_reduction_multi<1, K, kSpmvThreads>(a.m, d0, red_scratch);
}

// ------------------------------------------------------------------------------------------ K-wide vector passes
// Thread mapping of vec_loop<double>: packs of two consecutive rows in a fixed grid-stride order, an odd last row
// by thread 0 of CTA 0.  fr(row) handles one row (all K columns).
template <typename FR>
__device__ __forceinline__ void vec_loop_rows(long long n, FR fr) {
  constexpr int U = 2;
  const long long np = n / 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long ip = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; ip < np; ip += U * stride) {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (ip + u * stride < np) {
        fr(2 * (ip + u * stride));
        fr(2 * (ip + u * stride) + 1);
      }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = np * 2; i < n; ++i) fr(i);
}

// ConjugateGradient.h:43-67 for K columns
template <int K>
__global__ void __launch_bounds__(kVecThreads) cg_init_multi_kernel(const MultiArgs<K> a) {
  pdl_wait();  // no-op unless launched with programmatic stream serialization (B200S_PDL=1)
  __shared__ double scratch[32 * 3 * K];
  const bool guess = a.S[0].use_guess != 0;
  double v[3 * K];
#pragma unroll
  for (int j = 0; j < 3 * K; ++j) v[j] = 0.0;
  vec_loop_rows(a.n, [&](long long i) {
    double b[K], q[K], r[K], p[K];
    ldk<K>(a.b, i, b);
    const double d = a.invdiag[i];
    if (guess) ldk<K>(a.q, i, q);
#pragma unroll
    for (int j = 0; j < K; ++j) {
      r[j] = guess ? b[j] - q[j] : b[j];
      p[j] = d * r[j];
      v[3 * j + 0] = fma_rn(b[j], b[j], v[3 * j + 0]);
      v[3 * j + 1] = fma_rn(r[j], r[j], v[3 * j + 1]);
      v[3 * j + 2] = fma_rn(r[j], p[j], v[3 * j + 2]);
    }
    if (!guess) {
      double z[K];
#pragma unroll
      for (int j = 0; j < K; ++j) z[j] = 0.0;
      stk<K>(a.x, i, z);
    }
    stk<K>(a.r, i, r);
    stk<K>(a.p, i, p);
  });
  finish_reduction_multi<3, K, kVecThreads>(a, v, scratch);
}

// ConjugateGradient.h:75-84 for K columns (x update deferred as in cg_update_body)
template <int K>
__global__ void __launch_bounds__(kVecThreads) cg_update_multi_kernel(const MultiArgs<K> a) {
  pdl_wait();  // no-op unless launched with programmatic stream serialization (B200S_PDL=1)
  __shared__ double scratch[32 * 2 * K];
  if (multi_gated_out(a.S, a.gate)) return;
  double alpha[K];
  bool live[K];
#pragma unroll
  for (int j = 0; j < K; ++j) { alpha[j] = a.S[j].alpha; live[j] = a.S[j].stop == 0; }
  double v[2 * K];
#pragma unroll
  for (int j = 0; j < 2 * K; ++j) v[j] = 0.0;
  vec_loop_rows(a.n, [&](long long i) {
    double r[K], q[K];
    ldk<K>(a.r, i, r);
    ldk<K>(a.q, i, q);
    const double d = a.invdiag[i];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      if (live[j]) r[j] = fma_rn(-alpha[j], q[j], r[j]);
      const double z = d * r[j];
      v[2 * j + 0] = fma_rn(r[j], r[j], v[2 * j + 0]);
      v[2 * j + 1] = fma_rn(r[j], z, v[2 * j + 1]);
    }
    stk<K>(a.r, i, r);
  });
  finish_reduction_multi<2, K, kVecThreads>(a, v, scratch);
}

// ConjugateGradient.h:74 (deferred) and :81,:86 for K columns.  Per column: apply the pending x += alpha p; unless
// that column has just stopped, p = D^-1 r + beta p.
template <int K>
__global__ void __launch_bounds__(kVecThreads) cg_direction_multi_kernel(const MultiArgs<K> a, unsigned int* ticket) {
  pdl_wait();  // no-op unless launched with programmatic stream serialization (B200S_PDL=1)
  double alpha[K], beta[K];
  bool pend[K], upd[K];
  bool any = false;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    alpha[j] = a.S[j].alpha;
    beta[j] = a.S[j].beta;
    pend[j] = a.S[j].n_update != a.S[j].n_xapplied;
    upd[j] = pend[j] && a.S[j].stop == 0;
    any |= pend[j];
  }
  if (!any) return;
  vec_loop_rows(a.n, [&](long long i) {
    double x[K], p[K], r[K];
    ldk<K>(a.x, i, x);
    ldk<K>(a.p, i, p);
    ldk<K>(a.r, i, r);
    const double d = a.invdiag[i];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      if (pend[j]) x[j] = fma_rn(alpha[j], p[j], x[j]);
      if (upd[j]) p[j] = fma_rn(beta[j], p[j], d * r[j]);
    }
    stk<K>(a.x, i, x);
    stk<K>(a.p, i, p);
  });
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {
      *ticket = 0;
      for (int j = 0; j < K; ++j) a.S[j].n_xapplied = a.S[j].n_update;
    }
  }
}

// x = 0 for columns with ||b|| == 0, x = NaN for columns that ran into a non-finite residual (see finalize_kernel)
template <int K>
__global__ void __launch_bounds__(kVecThreads) finalize_multi_kernel(const MultiArgs<K> a) {
  pdl_wait();  // no-op unless launched with programmatic stream serialization (B200S_PDL=1)
  bool zero[K], nan[K], any = false;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    zero[j] = a.S[j].rhs_zero != 0;
    nan[j] = a.S[j].numerical_issue == 2;
    any |= zero[j] | nan[j];
  }
  if (!any) return;
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
#pragma unroll
    for (int j = 0; j < K; ++j)
      if (zero[j] || nan[j]) a.x[i * K + j] = zero[j] ? 0.0 : qnan;
}

// Column-major host layout (ld >= n) <-> interleaved device layout [n][K]; columns >= ncols are padded with zeros
// (a zero right-hand side stops at once: ConjugateGradient.h:46-52) and never copied back.
template <int K>
__global__ void interleave_kernel(long long n, int ncols, const double* __restrict__ src, long long ld,
                                  double* __restrict__ dst) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
#pragma unroll
    for (int j = 0; j < K; ++j) dst[i * K + j] = (j < ncols) ? src[j * ld + i] : 0.0;
}
template <int K>
__global__ void deinterleave_kernel(long long n, int ncols, const double* __restrict__ src, double* __restrict__ dst,
                                    long long ld) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
#pragma unroll
    for (int j = 0; j < K; ++j)
      if (j < ncols) dst[j * ld + i] = src[i * K + j];
}

}  // namespace b200s

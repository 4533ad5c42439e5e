// kernels.cuh -- hand-written sm_100a kernels of the sparse iterative-solve path.
//
//   spmv_staged_kernel   persistent CSR SpMV: matrix tiles streamed into shared memory by the bulk async-copy
//                        engine (cp.async.bulk + mbarrier, SASS UBLKCP), rows reduced from shared memory with a
//                        per-tile lanes-per-row choice, x gathered through L1/L2, optional fused dot products.
//   spmv_direct_kernel   plain sub-warp-per-row CSR SpMV from global memory (A/B baseline for the ncu evidence).
//   cg_* / bicg_*        the vector updates of ConjugateGradient.h:63-87 and BiCGSTAB.h:82-102, each fused with
//                        the dot products that follow it.
//   (halo push)          fused into the head of the SpMV kernels: boundary entries of x are stored into the
//                        neighbours' ghost slots over NVLink while the interior tiles are already streaming.
//
// Reductions are deterministic: every thread accumulates a fixed set of elements in a fixed order, warps are
// folded with shuffles, warps of a CTA in a fixed tree, CTAs by the last-arriving CTA in index order, ranks in
// rank order.  No floating-point atomics anywhere.
#pragma once
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>

#include "device_state.h"

namespace b200s {

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (TMA engine, no tensor map needed).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// Programmatic dependent launch: let the next kernel of the stream become resident early / wait until everything the
// previous kernel wrote is visible.  Both are no-ops for launches without the PDL attribute.  The trigger is issued
// LATE -- by each CTA when its streaming work is done, just before the reduction tail -- so that the next kernel's
// CTAs move in while the last CTA folds the partials and runs the scalar epilogue (and the SpMV's CTAs already
// prefetch their first tiles), but never squat on an SM whose current CTAs are still streaming (a trigger at kernel
// entry, round 1, cost 7 % at 256^3).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_u32(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_f64(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// ---------------------------------------------------------------------------------- rounding-explicit arithmetic
// The reference's SpMV row loop rounds each product and each add separately (see oracle/oracle_body.h); the other
// updates are contracted to FMAs.  Both are spelled out so that nvcc's own contraction cannot change them.
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// ------------------------------------------------------------------------------------------- block reductions
template <int NV>
__device__ __forceinline__ void warp_reduce(double (&v)[NV]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
}

// All threads call; result valid in thread 0.  THREADS is a multiple of 32, <= 1024.
template <int NV, int THREADS>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double* scratch /* [32*NV] shared */) {
  warp_reduce<NV>(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // scratch may still be in use by a previous call
  if (lane == 0)
#pragma unroll
    for (int j = 0; j < NV; ++j) scratch[warp * NV + j] = v[j];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = (lane < THREADS / 32) ? scratch[lane * NV + j] : 0.0;
    warp_reduce<NV>(v);
  }
}

// ------------------------------------------------------------------------------ cross-rank all-reduce (last CTA)
// One-shot all-gather of the per-rank partials into every rank's mailbox over NVLink peer stores, then a sum in
// rank order, so that every rank obtains bit-identical scalars and takes identical branches.
// Low-latency protocol: every 8-byte word carries 32 bits of payload and the 32-bit sequence number of the
// reduction, so a word is its own arrival flag -- no fence and no separate flag store on the critical path
// (a fence + release-store design measured 13 us per reduction on 4 GPUs; fenceless words cost one NVLink hop).
// Two mailbox parities: a rank can be at most one reduction ahead of the slowest reader, because it needs that
// reader's next partial to advance further.  Called by ALL threads of the last CTA; values in/out in thread 0.
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

constexpr int kBoxWords = 8;  // per (parity, source rank): up to 4 doubles as 8 tagged words

// Every spin on another CTA's or another GPU's progress is bounded: when the limit expires (a peer died on the host
// before launching, or ranks were given diverging inputs) the waiter raises comm_error + stop, the loop winds down
// through its normal gates, and the host reports B200S_ERR_COMM instead of hanging inside a graph.
__device__ __forceinline__ unsigned long long spin_limit_ns(const Scalars* S) {
  const unsigned long long t = S->comm_timeout_ns;
  return t ? t : 20000000000ull;
}
__device__ __forceinline__ bool spin_expired(Scalars* S, unsigned long long t0, unsigned& spins) {
  if ((++spins & 0x3ffu) != 0) return false;
  if (globaltimer_ns() - t0 <= spin_limit_ns(S)) return false;
  S->comm_error = 1;
  S->stop = 1;
  __threadfence();
  return true;
}

__device__ __forceinline__ void allreduce_ranks(const CommDev& c, Scalars* S, double* v, int n) {
  __shared__ unsigned s_seq;
  __shared__ double s_in[4];
  __shared__ unsigned s_words[kMaxWorld * kBoxWords];
  const int tid = threadIdx.x;
  unsigned long long t0 = 0;
  if (tid == 0) {
    s_seq = ++S->red_seq;
    for (int j = 0; j < n; ++j) s_in[j] = v[j];
    if (c.world > 1) t0 = globaltimer_ns();
  }
  if (c.world <= 1) return;
  __syncthreads();
  const unsigned seq = s_seq;
  const int par = seq & 1;
  const int nw = 2 * n;  // words per rank
  if (tid < c.world * nw) {
    const int peer = tid / nw, w = tid - peer * nw;
    const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(s_in[w >> 1]));
    const unsigned long long payload = (w & 1) ? (bits >> 32) : (bits & 0xffffffffull);
    unsigned long long* box = reinterpret_cast<unsigned long long*>((peer == c.rank) ? c.box_self : c.box_peer[peer]);
    st_relaxed_sys_u64(box + (par * kMaxWorld + c.rank) * kBoxWords + w,
                       (static_cast<unsigned long long>(seq) << 32) | payload);
    // the same thread now waits for word w of source rank `peer`
    const unsigned long long* in =
        reinterpret_cast<const unsigned long long*>(c.box_self) + (par * kMaxWorld + peer) * kBoxWords + w;
    unsigned long long got;
    unsigned spins = 0;
    const unsigned long long tw = globaltimer_ns();
    do {
      got = ld_relaxed_sys_u64(in);
    } while (static_cast<unsigned>(got >> 32) != seq && !spin_expired(S, tw, spins));
    s_words[peer * kBoxWords + w] = static_cast<unsigned>(got);
  }
  __syncthreads();
  if (tid == 0) {
    for (int j = 0; j < n; ++j) {
      double sum = 0.0;
      for (int src = 0; src < c.world; ++src) {  // rank order: identical on every rank
        const unsigned long long bits = (static_cast<unsigned long long>(s_words[src * kBoxWords + 2 * j + 1]) << 32) |
                                        s_words[src * kBoxWords + 2 * j];
        sum += __longlong_as_double(static_cast<long long>(bits));
      }
      v[j] = sum;
    }
    S->t_allreduce += globaltimer_ns() - t0;
  }
}

// --------------------------------------------------------------------------------- scalar logic of the solvers
// Runs in exactly one thread per reduction, after the values are reduced over CTAs and ranks.  This is the
// reference's host-side control flow moved onto the device (citations per case).
// RealScalar arithmetic of the reference for both instantiations: in float mode (ctx.f32) every scalar operation
// is a float operation -- performed here in double and rounded to float after each step, which gives the same
// result (products of two floats are exact in double; for + - / sqrt, 53 >= 2*24+2 bits make the double rounding
// innocuous).  Reduced dot products are rounded to float once.
__device__ __forceinline__ double rs(bool f32, double x) { return f32 ? static_cast<double>(static_cast<float>(x)) : x; }

__device__ inline void run_epilogue(const RedCtx& ctx, const double* vin, double* history) {
  Scalars* S = ctx.S;
  const bool f = ctx.f32 != 0;
  double v[4];
  for (int j = 0; j < 4; ++j) v[j] = rs(f, vin[j]);
  const double real_min = f ? static_cast<double>(FLT_MIN) : DBL_MIN;
  const double real_eps = f ? static_cast<double>(FLT_EPSILON) : DBL_EPSILON;
  switch (ctx.epilogue) {
    case kEpiCgInit: {  // ConjugateGradient.h:45-67
      const double bb = v[0], rr = v[1], rz = v[2];
      S->bb = bb; S->rr = rr; S->iter = 0; S->converged = 0; S->rhs_zero = 0; S->stop = 0; S->numerical_issue = 0;
      S->hist_len = 0; S->spmv_count = S->use_guess ? 1 : 0; S->n_update = 0; S->n_xapplied = 0;
      if (bb == 0.0) { S->rhs_zero = 1; S->stop = 1; S->rr = 0.0; break; }   // :46-52 (error = 0)
      double thr = rs(f, rs(f, S->tol * S->tol) * bb);                        // :53-54
      if (thr < real_min) thr = real_min;
      S->thr = thr;
      if (rr < thr) { S->stop = 1; S->converged = 1; break; }                 // :56-61
      S->abs_new = rz;                                                        // :67
      // A non-finite norm can never pass `rr < thr`: the reference would spin through all maxIters iterations on NaNs
      // and return NoConvergence with iters = maxIters, error = NaN and an all-NaN x (one iteration spreads the NaN
      // through alpha).  The loop stops here instead and the host reports exactly those outputs.
      if (!(rr == rr) || !(bb == bb)) { S->numerical_issue = (S->max_iters >= 1) ? 2 : 1; S->stop = 1; }
      if (S->max_iters <= 0) S->stop = 1;                                     // while(i < maxIters) never entered
    } break;
    case kEpiCgPAp: {  // :73
      S->pAp = v[0];
      S->alpha = rs(f, S->abs_new / v[0]);
      S->spmv_count++;
    } break;
    case kEpiCgUpdate: {  // :77-87
      const double rr = v[0], rz = v[1];
      S->rr = rr;
      S->n_update++;  // x += alpha p of this iteration is still owed (applied by the direction pass)
      if (history && S->hist_len < kHistoryCap) history[S->hist_len++] = rr;
      if (rr < S->thr) { S->stop = 1; S->converged = 1; break; }  // :78-79 break before i++
      if (!(rr == rr)) {  // see kEpiCgInit: two more reference iterations would turn all of x into NaN
        S->numerical_issue = (S->max_iters - S->iter >= 2) ? 2 : 1;
        S->stop = 1;
        break;
      }
      S->abs_old = S->abs_new;
      S->abs_new = rz;                        // :84
      S->beta = rs(f, rz / S->abs_old);       // :85
      S->iter++;                              // :87
      if (S->iter >= S->max_iters) S->stop = 1;
    } break;
    case kEpiBiInit: {  // BiCGSTAB.h:45-65
      const double bb = v[0], rr = v[1];
      S->bb = bb; S->rr = rr; S->r0_sqnorm = rr; S->iter = 0; S->restarts = 0; S->converged = 0; S->rhs_zero = 0;
      S->stop = 0; S->numerical_issue = 0; S->restart = 0; S->hist_len = 0; S->spmv_count = 1;
      if (bb == 0.0) { S->rhs_zero = 1; S->stop = 1; break; }  // :47-51 (iters / tol_error untouched)
      S->rho_old = 1.0; S->alpha = 1.0; S->w = 1.0;            // :52-54
      S->thr = rs(f, rs(f, S->tol * S->tol) * bb);             // :62
      S->eps2 = rs(f, real_eps * real_eps);                    // :63
      if (!(rr > S->thr && 0 < S->max_iters)) { S->stop = 1; S->converged = !(rr > S->thr); break; }  // :67
      S->rho = rr;                                             // :71 r0.dot(r) with r0 == r
      S->restart = (fabs(S->rho) < rs(f, S->eps2 * S->r0_sqnorm)) ? 1 : 0;  // :72
    } break;
    case kEpiBiR0V: {  // :89
      S->r0v = v[0];
      S->alpha = rs(f, S->rho / v[0]);
      S->spmv_count++;
    } break;
    case kEpiBiTsTt: {  // :95-99
      S->ts = v[0]; S->tt = v[1];
      S->w = (v[1] > 0.0) ? rs(f, v[0] / v[1]) : 0.0;
      S->spmv_count++;
    } break;
    case kEpiBiUpdate: {  // :100-102 then the loop head :67-72 of the next iteration
      const double rr = v[0], rho_next = v[1];
      S->rr = rr;
      S->iter++;
      if (history && S->hist_len < kHistoryCap) history[S->hist_len++] = rr;
      if (!(rr > S->thr && S->iter < S->max_iters)) { S->stop = 1; S->converged = !(rr > S->thr); break; }
      S->rho_old = S->rho;
      S->rho = rho_next;
      S->restart = (fabs(S->rho) < rs(f, S->eps2 * S->r0_sqnorm)) ? 1 : 0;
    } break;
    case kEpiBiRestart: {  // :76-80
      S->rho = S->r0_sqnorm = v[0];
      if (S->restarts++ == 0) S->iter = 0;
      S->restart = 0;
      S->spmv_count++;
    } break;
    default: break;
  }
}

__device__ __forceinline__ bool gated_out(const Scalars* S, int gate) {
  switch (gate) {
    case kGateLoop: return S->stop != 0;
    case kGateRestart: return S->stop != 0 || S->restart == 0;
    case kGateGuess: return S->use_guess == 0;
    default: return false;
  }
}

// Device-side timeline: time between consecutive reduction epilogues, accumulated per epilogue kind.
__device__ __forceinline__ void timeline_mark(Scalars* S, int e) {
  const unsigned long long now = globaltimer_ns();
  if (e == kEpiCgInit || e == kEpiBiInit) {
    for (int i = 0; i < 12; ++i) { S->t_phase[i] = 0; S->n_phase[i] = 0; }
    S->t_first = now;
    S->t_allreduce = 0;
    S->t_halo_wait = 0;
  } else if (e > 0 && e < 12) {
    S->t_phase[e] += now - S->t_last;
    S->n_phase[e]++;
  }
  S->t_last = now;
  S->t_end = now;
}

// Final stage of every reduction: publish this CTA's partials, take a ticket; the last CTA folds all partials in
// CTA order, all-reduces over ranks, runs the scalar epilogue and (in WHILE-graph mode) sets the loop condition.
template <int NV, int THREADS>
__device__ __forceinline__ void finish_reduction(const RedCtx& ctx, double (&v)[NV], double* scratch,
                                                 double* history) {
  __shared__ int s_last;
  block_reduce<NV, THREADS>(v, scratch);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int j = 0; j < NV; ++j) ctx.partials[blockIdx.x * 4 + j] = v[j];
    __threadfence();
    const unsigned ticket = atomicAdd(ctx.counter, 1u);
    s_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double t[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) t[j] = 0.0;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += THREADS)
#pragma unroll
    for (int j = 0; j < NV; ++j) t[j] += __ldcg(ctx.partials + b * 4 + j);
  block_reduce<NV, THREADS>(t, scratch);
  double r[4] = {0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < NV; ++j) r[j] = t[j];
  allreduce_ranks(ctx.comm, ctx.S, r, NV);
  if (threadIdx.x == 0) {
    *ctx.counter = 0;
    if (ctx.bump_halo) ctx.S->halo_seq++;  // this kernel carried a halo exchange: retire its sequence number
    timeline_mark(ctx.S, ctx.epilogue);
    run_epilogue(ctx, r, history);
    if (ctx.S->comm_error) ctx.S->stop = 1;  // an expired spin ends the loop whatever the epilogue decided
    if (ctx.set_cond) cudaGraphSetConditional(static_cast<cudaGraphConditionalHandle>(ctx.cond_handle), ctx.S->stop ? 0u : 1u);
  }
}

// ---------------------------------------------------------------------------------------------------- halo push
// Send plan of one rank: which owned entries go to which peer's ghost slots (pointers are NVLink-mapped peer memory).
template <typename T>
struct HaloArgs {
  int enabled;                // world > 1
  int npush;                  // CTAs [0, npush) carry the push; the others go straight to their tiles
  const int32_t* send_rows;   // local rows to send, grouped by destination
  long long send_offsets[kMaxWorld];
  long long send_counts[kMaxWorld];
  T* dst[kMaxWorld];          // peer q's ghost slots for my entries
  unsigned int* counter;      // CTA ticket for "all my stores are out"
};

// Every CTA stores its share of the boundary entries straight into the peers' ghost slots, fences, and takes a
// ticket; the last CTA publishes the new halo sequence number to the peers.  The sequence number is S->halo_seq + 1,
// where S->halo_seq is only advanced by the kernel's final single-thread epilogue, i.e. it is stable while any CTA of
// this kernel (pusher or waiter) reads it.
template <typename T>
__device__ __forceinline__ void halo_push(const HaloArgs<T>& hl, const CommDev& c, const Scalars* S, const T* x) {
  if (static_cast<int>(blockIdx.x) >= hl.npush) return;
  for (int q = 0; q < c.world; ++q) {
    const long long n = hl.send_counts[q];
    if (n == 0) continue;
    const int32_t* rows = hl.send_rows + hl.send_offsets[q];
    T* dst = hl.dst[q];
    for (long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; k < n;
         k += static_cast<long long>(hl.npush) * blockDim.x)
      dst[k] = x[rows[k]];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(hl.counter, 1u);
    if (ticket == static_cast<unsigned>(hl.npush) - 1) {
      *hl.counter = 0;
      __threadfence_system();
      const unsigned seq = S->halo_seq + 1;
      for (int q = 0; q < c.world; ++q)
        if (hl.send_counts[q] > 0) st_relaxed_sys_u32(c.halo_flag_peer[q] + c.rank, seq);  // ordered by the fence above
    }
  }
}

__device__ __forceinline__ void halo_wait(const CommDev& c, Scalars* S, unsigned recv_mask) {
  const unsigned want = S->halo_seq + 1;
  unsigned spins = 0;
  const unsigned long long tw = globaltimer_ns();
  for (int src = 0; src < c.world; ++src)
    if (recv_mask & (1u << src))
      while (static_cast<int>(ld_acquire_sys(c.halo_flag_self + src) - want) < 0)
        if (spin_expired(S, tw, spins)) return;
}

// ------------------------------------------------------------------------------------------- staged CSR SpMV
template <typename T>
struct SpmvArgs {
  const Tile* tiles;
  int ntiles;
  int first_boundary_tile;  // tiles [first_boundary_tile, ntiles) gather ghost entries
  const int32_t* rowptr;
  const int32_t* colidx;
  const T* vals;
  const T* x;       // extended vector [owned | ghost]
  T* y;
  const T* w;       // left operand of the first fused dot (nullptr: use x)
  int stages, cap_nnz, cap_rows;
  int tail_blk;     // float only: reference rounding pattern (0 = every product rounded)
  int evict_first;  // stream the matrix through L2 with an evict-first policy
  unsigned recv_mask;  // ranks whose halo must have arrived before boundary tiles (multi-GPU)
  HaloArgs<T> halo;    // this rank's pushes, issued at the head of the kernel
  double* history;
  RedCtx red;
};

template <typename T>
__host__ __device__ inline size_t spmv_stage_bytes(int cap_nnz, int cap_rows) {
  size_t v = (static_cast<size_t>(cap_nnz) + 16) * sizeof(T);
  size_t c = (static_cast<size_t>(cap_nnz) + 8) * 4;
  size_t r = (static_cast<size_t>(cap_rows) + 1 + 8) * 4;
  auto up = [](size_t b) { return (b + 127) & ~static_cast<size_t>(127); };
  return up(v) + up(c) + up(r);
}

// x gather: through the non-coherent path when x is read-only for the whole kernel (one product per launch), or as
// an ordinary L1-cached load when other phases of the same (persistent) kernel rewrite x between products.
template <bool NC, typename T>
__device__ __forceinline__ T ldx(const T* p) {
  if (NC) return __ldg(p);
  return *p;
}

template <typename T, int LG, bool STREAM, bool NC>
__device__ __forceinline__ void tile_rows_reduce(const T* __restrict__ sv, const int32_t* __restrict__ sc,
                                                 const int32_t* __restrict__ srp, int nrows, int row0, int vb0,
                                                 int cb0, const T* x, T* __restrict__ y, const T* w, int tail_blk,
                                                 double& d0, double& d1) {
  constexpr int L = 1 << LG;
  const int lane = threadIdx.x & (L - 1);
  const int grp = threadIdx.x >> LG;
  constexpr int NGRP = kSpmvThreads >> LG;
  for (int base = 0; base < nrows; base += NGRP) {
    const int r = base + grp;
    T sum = T(0);
    if (r < nrows) {
      int k = srp[r];
      const int k1 = srp[r + 1];
      if (STREAM) {
        for (k += lane; k < k1; k += L) sum = add_rn(sum, sv[k - vb0]);
      } else {
        // Batches of B entries per lane with every shared-memory read and every x gather issued before the first
        // add: one memory latency per batch instead of one per entry.  Slots past the end of the row multiply
        // 0 * 0 and add +0, which leaves the sum unchanged (the final -0 -> +0 normalisation happens anyway).
        constexpr int B = (L == 1) ? 8 : 4;
        const int kfuse = (sizeof(T) == 4 && L == 1 && tail_blk > 0) ? k + ((k1 - k) / tail_blk) * tail_blk : k1;
        for (k += lane; k < k1; k += B * L) {
          int c[B];
          T v[B], xv[B];
#pragma unroll
          for (int j = 0; j < B; ++j) {
            const bool in = (k + j * L) < k1;
            c[j] = in ? sc[k + j * L - cb0] : 0;
            v[j] = in ? sv[k + j * L - vb0] : T(0);
          }
#pragma unroll
          for (int j = 0; j < B; ++j) xv[j] = ((k + j * L) < k1) ? ldx<NC>(x + c[j]) : T(0);
#pragma unroll
          for (int j = 0; j < B; ++j) {
            if (sizeof(T) == 4 && L == 1 && (k + j) >= kfuse)
              sum = fma_rn(v[j], xv[j], sum);  // float: the reference's fused scalar epilogue
            else
              sum = add_rn(sum, mul_rn(v[j], xv[j]));
          }
        }
      }
    }
    if (L > 1) {
#pragma unroll
      for (int o = L / 2; o > 0; o >>= 1) sum = add_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));
    }
    if (r < nrows && lane == 0) {
      sum = add_rn(sum, T(0));  // -0 -> +0, as `res += alpha*tmp` on a zeroed destination does
      y[row0 + r] = sum;
      if (w) d0 = fma_rn(static_cast<double>(w[row0 + r]), static_cast<double>(sum), d0);
      d1 = fma_rn(static_cast<double>(sum), static_cast<double>(sum), d1);
    }
  }
}

// Phase 2 of the two-phase ("stream") tiles: the products already sit in shared memory, each row only has to be
// summed.  Kernel selection at ROW granularity from the row length: a short row is summed by one thread (scalar), a
// longer one by a whole warp (32 lanes + shuffle tree) -- in a tile of 250 short rows and one row of 1,500 entries
// nobody waits for a single thread walking the long row.  Fixed row-to-thread mapping: deterministic.
template <typename T>
__device__ __forceinline__ void tile_rows_reduce_binned(const T* __restrict__ sv, const int32_t* __restrict__ srp,
                                                        int nrows, int row0, int vb0, T* __restrict__ y, const T* w,
                                                        double& d0, double& d1) {
  constexpr int kShortRow = 12;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto finish = [&](int r, T sum) {
    sum = add_rn(sum, T(0));  // -0 -> +0
    y[row0 + r] = sum;
    if (w) d0 = fma_rn(static_cast<double>(w[row0 + r]), static_cast<double>(sum), d0);
    d1 = fma_rn(static_cast<double>(sum), static_cast<double>(sum), d1);
  };
  for (int r = tid; r < nrows; r += kSpmvThreads) {  // scalar bin
    const int k0 = srp[r], k1 = srp[r + 1];
    if (k1 - k0 <= kShortRow) {
      T sum = T(0);
      for (int k = k0; k < k1; ++k) sum = add_rn(sum, sv[k - vb0]);
      finish(r, sum);
    }
  }
  for (int r = warp; r < nrows; r += kSpmvThreads / 32) {  // warp bin
    const int k0 = srp[r], k1 = srp[r + 1];
    if (k1 - k0 > kShortRow) {
      T sum = T(0);
      for (int k = k0 + lane; k < k1; k += 32) sum = add_rn(sum, sv[k - vb0]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum = add_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));
      if (lane == 0) finish(r, sum);
    }
  }
}

// Per-CTA state of the staged product.  `seq` counts the tiles this CTA has consumed since the mbarriers were
// initialised, which fixes the stage (seq % stages) and the barrier parity ((seq / stages) & 1) of every tile, also
// across successive products inside one persistent kernel.
template <typename T>
struct SpmvCta {
  unsigned char* smem;
  uint64_t* full_bar;
  T* long_scratch;
  size_t stage_bytes, v_bytes, c_bytes;
  uint64_t policy;
  unsigned seq;
};

template <typename T>
__device__ __forceinline__ void spmv_cta_init(const SpmvArgs<T>& a, SpmvCta<T>& cx, unsigned char* smem,
                                              uint64_t* full_bar, T* long_scratch) {
  cx.smem = smem;
  cx.full_bar = full_bar;
  cx.long_scratch = long_scratch;
  cx.stage_bytes = spmv_stage_bytes<T>(a.cap_nnz, a.cap_rows);
  cx.v_bytes = ((static_cast<size_t>(a.cap_nnz) + 16) * sizeof(T) + 127) & ~static_cast<size_t>(127);
  cx.c_bytes = ((static_cast<size_t>(a.cap_nnz) + 8) * 4 + 127) & ~static_cast<size_t>(127);
  cx.seq = 0;
  cx.policy = a.evict_first ? policy_evict_first() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
}

// One elected thread: start the bulk copies of tile t into stage s (values, column indices, row-pointer slice).
template <typename T>
__device__ __forceinline__ void spmv_issue_tile(const SpmvArgs<T>& a, const SpmvCta<T>& cx, const Tile tl, int s);
template <typename T>
__device__ __forceinline__ void spmv_issue(const SpmvArgs<T>& a, const SpmvCta<T>& cx, int t, int s) {
  spmv_issue_tile<T>(a, cx, a.tiles[t], s);
}
template <typename T>
__device__ __forceinline__ void spmv_issue_tile(const SpmvArgs<T>& a, const SpmvCta<T>& cx, const Tile tl, int s) {
  constexpr int VA = 16 / sizeof(T);  // elements per 16-byte unit of the value array
  uint64_t* bar = &cx.full_bar[s];
  if ((tl.meta >> 24) & kTileLong) { mbar_arrive(bar); return; }
  const int nrows = tl.meta & 0xFFFF;
  const int vb0 = tl.nnz0 & ~(VA - 1), vb1 = (tl.nnz0 + tl.nnz + VA - 1) & ~(VA - 1);
  const int cb0 = tl.nnz0 & ~3, cb1 = (tl.nnz0 + tl.nnz + 3) & ~3;
  const int rb0 = tl.row0 & ~3, rb1 = (tl.row0 + nrows + 1 + 3) & ~3;
  const unsigned vby = static_cast<unsigned>(vb1 - vb0) * sizeof(T), cby = static_cast<unsigned>(cb1 - cb0) * 4u,
                 rby = static_cast<unsigned>(rb1 - rb0) * 4u;
  unsigned char* st = cx.smem + static_cast<size_t>(s) * cx.stage_bytes;
  mbar_expect_tx(bar, vby + cby + rby);
  if (a.evict_first) {
    if (vby) bulk_g2s_hint(st, a.vals + vb0, vby, bar, cx.policy);
    if (cby) bulk_g2s_hint(st + cx.v_bytes, a.colidx + cb0, cby, bar, cx.policy);
    bulk_g2s_hint(st + cx.v_bytes + cx.c_bytes, a.rowptr + rb0, rby, bar, cx.policy);
  } else {
    if (vby) bulk_g2s(st, a.vals + vb0, vby, bar);
    if (cby) bulk_g2s(st + cx.v_bytes, a.colidx + cb0, cby, bar);
    bulk_g2s(st + cx.v_bytes + cx.c_bytes, a.rowptr + rb0, rby, bar);
  }
}

// Fill the ring with this CTA's first tiles (independent of x: may be issued long before the product starts).
template <typename T>
__device__ __forceinline__ void spmv_prefetch(const SpmvArgs<T>& a, const SpmvCta<T>& cx) {
  if (threadIdx.x == 0)
    for (int k = 0; k < a.stages; ++k) {
      const int t = blockIdx.x + k * gridDim.x;
      if (t < a.ntiles) spmv_issue(a, cx, t, (cx.seq + k) % a.stages);
    }
}

// This CTA's share of one product y = A x (+ partial dots).  spmv_prefetch must have been called for this round.
template <typename T, int NDOT, bool NC>
__device__ __forceinline__ void spmv_tiles(const SpmvArgs<T>& a, SpmvCta<T>& cx, double& d0, double& d1) {
  constexpr int VA = 16 / sizeof(T);
  const int tid = threadIdx.x;
  const int S = a.stages;
  const int G = gridDim.x;
  const T* w = (NDOT >= 1) ? (a.w ? a.w : a.x) : nullptr;
  bool halo_ready = (a.recv_mask == 0);
  int k = 0;
  // Tile descriptors are fetched one round ahead (and, by the issuing thread, S rounds ahead): a descriptor load is
  // an L2 round trip that would otherwise sit in front of every tile -- a larger share of a float tile, which drains
  // in two thirds of the time of a double tile.
  Tile tl_next = (static_cast<int>(blockIdx.x) < a.ntiles) ? a.tiles[blockIdx.x] : Tile{};
  for (;; ++k) {
    const int t = blockIdx.x + k * G;
    if (t >= a.ntiles) break;
    const unsigned q = cx.seq + k;
    const int s = q % S;
    const unsigned parity = (q / S) & 1;
    const Tile tl = tl_next;
    if (t + G < a.ntiles) tl_next = a.tiles[t + G];
    const bool refill = (tid == 0) && (t + S * G < a.ntiles);
    Tile tl_refill = tl;
    if (refill) tl_refill = a.tiles[t + S * G];
    const int flags = (tl.meta >> 24) & 0xFF;
    const int nrows = tl.meta & 0xFFFF;
    const int lg = (tl.meta >> 16) & 0xFF;

    if (!halo_ready && t >= a.first_boundary_tile) {  // ghost entries must have landed (multi-GPU)
      if (tid == 0) {
        const unsigned long long t0 = (blockIdx.x == 0) ? globaltimer_ns() : 0ull;
        halo_wait(a.red.comm, a.red.S, a.recv_mask);
        if (blockIdx.x == 0) a.red.S->t_halo_wait += globaltimer_ns() - t0;
      }
      __syncthreads();
      halo_ready = true;
    }

    mbar_wait(&cx.full_bar[s], parity);

    if (flags & kTileLong) {
      // one row longer than a stage: the whole CTA streams it from global memory, coalesced
      T sum = T(0);
      const int k1 = tl.nnz0 + tl.nnz;
      for (int kk = tl.nnz0 + tid; kk < k1; kk += kSpmvThreads)
        sum = add_rn(sum, mul_rn(a.vals[kk], ldx<NC>(a.x + a.colidx[kk])));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum = add_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));
      __syncthreads();
      if ((tid & 31) == 0) cx.long_scratch[tid >> 5] = sum;
      __syncthreads();
      if (tid < 32) {
        sum = (tid < kSpmvThreads / 32) ? cx.long_scratch[tid] : T(0);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum = add_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));
        if (tid == 0) {
          sum = add_rn(sum, T(0));
          a.y[tl.row0] = sum;
          if (w) d0 = fma_rn(static_cast<double>(w[tl.row0]), static_cast<double>(sum), d0);
          d1 = fma_rn(static_cast<double>(sum), static_cast<double>(sum), d1);
        }
      }
    } else {
      unsigned char* st = cx.smem + static_cast<size_t>(s) * cx.stage_bytes;
      T* sv = reinterpret_cast<T*>(st);
      const int32_t* sc = reinterpret_cast<const int32_t*>(st + cx.v_bytes);
      const int32_t* srp = reinterpret_cast<const int32_t*>(st + cx.v_bytes + cx.c_bytes) + (tl.row0 - (tl.row0 & ~3));
      const int vb0 = tl.nnz0 & ~(VA - 1);
      const int cb0 = tl.nnz0 & ~3;
      if (flags & kTileStream) {
        // phase 1: products, perfectly balanced over the CTA (CSR-stream); phase 2 sums them per row
        // All shared-memory reads and all gathers of a thread are issued before its first store: the stores into sv
        // could alias the index reads as far as the compiler knows, which would serialise the loop to one gather in
        // flight per thread.  (A tile holds at most 8 * 256 entries by default; larger tiles take more rounds.)
        const int k1 = tl.nnz0 + tl.nnz;
        for (int k0 = tl.nnz0 + tid; k0 < k1; k0 += 8 * kSpmvThreads) {
          int c[8];
          T v[8], xv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int kk = k0 + j * kSpmvThreads;
            c[j] = (kk < k1) ? sc[kk - cb0] : 0;
            v[j] = (kk < k1) ? sv[kk - vb0] : T(0);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) xv[j] = (k0 + j * kSpmvThreads < k1) ? ldx<NC>(a.x + c[j]) : T(0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int kk = k0 + j * kSpmvThreads;
            if (kk < k1) sv[kk - vb0] = mul_rn(v[j], xv[j]);
          }
        }
        __syncthreads();
        tile_rows_reduce_binned<T>(sv, srp, nrows, tl.row0, vb0, a.y, w, d0, d1);
      } else {
        switch (lg) {
          case 0: tile_rows_reduce<T, 0, false, NC>(sv, sc, srp, nrows, tl.row0, vb0, cb0, a.x, a.y, w, a.tail_blk, d0, d1); break;
          case 1: tile_rows_reduce<T, 1, false, NC>(sv, sc, srp, nrows, tl.row0, vb0, cb0, a.x, a.y, w, 0, d0, d1); break;
          case 2: tile_rows_reduce<T, 2, false, NC>(sv, sc, srp, nrows, tl.row0, vb0, cb0, a.x, a.y, w, 0, d0, d1); break;
          case 3: tile_rows_reduce<T, 3, false, NC>(sv, sc, srp, nrows, tl.row0, vb0, cb0, a.x, a.y, w, 0, d0, d1); break;
          case 4: tile_rows_reduce<T, 4, false, NC>(sv, sc, srp, nrows, tl.row0, vb0, cb0, a.x, a.y, w, 0, d0, d1); break;
          default: tile_rows_reduce<T, 5, false, NC>(sv, sc, srp, nrows, tl.row0, vb0, cb0, a.x, a.y, w, 0, d0, d1); break;
        }
      }
    }
    __syncthreads();  // every thread is done with stage s
    if (refill) spmv_issue_tile<T>(a, cx, tl_refill, s);
  }
  cx.seq += k;
}

template <typename T, int NDOT>
__global__ void __launch_bounds__(kSpmvThreads) spmv_staged_kernel(const SpmvArgs<T> a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[8];
  __shared__ double red_scratch[32 * 2];
  __shared__ T long_scratch[32];
  // Prologue that does not depend on the previous kernel: barrier setup and the first tile copies (matrix data is
  // constant).  It overlaps the predecessor's tail under programmatic dependent launch.
  SpmvCta<T> cx;
  spmv_cta_init(a, cx, smem, full_bar, long_scratch);
  spmv_prefetch(a, cx);
  pdl_wait();
  if (gated_out(a.red.S, a.red.gate)) {  // loop already stopped: let the copies land, then leave
    for (int k = 0; k < a.stages; ++k)
      if (static_cast<int>(blockIdx.x + k * gridDim.x) < a.ntiles) mbar_wait(&full_bar[k], 0);
    return;
  }
  // multi-GPU: my boundary entries go out to the neighbours while the first tiles are in flight
  if (a.halo.enabled) halo_push<T>(a.halo, a.red.comm, a.red.S, a.x);

  double d0 = 0.0, d1 = 0.0;
  spmv_tiles<T, NDOT, true>(a, cx, d0, d1);

  if (a.red.epilogue != kEpiNone) {
    if (NDOT == 0) {
      double v[1] = {0.0};
      pdl_launch_dependents();
      finish_reduction<1, kSpmvThreads>(a.red, v, red_scratch, a.history);
    } else if (NDOT == 1) {
      double v[1] = {d0};
      pdl_launch_dependents();
      finish_reduction<1, kSpmvThreads>(a.red, v, red_scratch, a.history);
    } else {
      double v[2] = {d0, d1};
      pdl_launch_dependents();
      finish_reduction<2, kSpmvThreads>(a.red, v, red_scratch, a.history);
    }
  }
}

// ------------------------------------------------------------------------------------------- direct CSR SpMV
// L = 2^LG lanes per row, straight from global memory; rows are dealt to lane groups in a fixed grid-stride order.
template <typename T, int LG, int NDOT>
__global__ void __launch_bounds__(kSpmvThreads) spmv_direct_kernel(const SpmvArgs<T> a, int rows) {
  __shared__ double red_scratch[32 * 2];
  pdl_wait();
  if (gated_out(a.red.S, a.red.gate)) return;
  constexpr int L = 1 << LG;
  const int lane = threadIdx.x & (L - 1);
  const long long grp0 = (static_cast<long long>(blockIdx.x) * kSpmvThreads + threadIdx.x) >> LG;
  const long long ngrp = (static_cast<long long>(gridDim.x) * kSpmvThreads) >> LG;
  const T* w = (NDOT >= 1) ? (a.w ? a.w : a.x) : nullptr;
  double d0 = 0.0, d1 = 0.0;
  if (a.halo.enabled) halo_push<T>(a.halo, a.red.comm, a.red.S, a.x);
  if (a.recv_mask != 0) {  // direct kernel has no interior/boundary split: wait for the halo up front
    if (threadIdx.x == 0) halo_wait(a.red.comm, a.red.S, a.recv_mask);
    __syncthreads();
  }
  for (long long base = 0; base < rows; base += ngrp) {
    const long long r = base + grp0;
    T sum = T(0);
    if (r < rows) {
      const int k1 = a.rowptr[r + 1];
      for (int k = a.rowptr[r] + lane; k < k1; k += L) sum = add_rn(sum, mul_rn(a.vals[k], __ldg(a.x + a.colidx[k])));
    }
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) sum = add_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));
    if (r < rows && lane == 0) {
      sum = add_rn(sum, T(0));
      a.y[r] = sum;
      if (w) d0 = fma_rn(static_cast<double>(w[r]), static_cast<double>(sum), d0);
      d1 = fma_rn(static_cast<double>(sum), static_cast<double>(sum), d1);
    }
  }
  if (a.red.epilogue != kEpiNone) {
    if (NDOT <= 1) {
      double v[1] = {NDOT ? d0 : 0.0};
      pdl_launch_dependents();
      finish_reduction<1, kSpmvThreads>(a.red, v, red_scratch, a.history);
    } else {
      double v[2] = {d0, d1};
      pdl_launch_dependents();
      finish_reduction<2, kSpmvThreads>(a.red, v, red_scratch, a.history);
    }
  }
}

// ------------------------------------------------------------------------------------------------ Jacobi setup
// BasicPreconditioners.h:64-79: first stored entry with inner index == j; missing or zero -> 1.
template <typename T>
__global__ void jacobi_factorize_kernel(int rows, const int32_t* __restrict__ rowptr,
                                        const int32_t* __restrict__ colidx, const T* __restrict__ vals,
                                        T* __restrict__ invdiag, int identity) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= rows) return;
  T d = T(1);
  if (!identity) {
    int k = rowptr[j];
    const int e = rowptr[j + 1];
    while (k < e && colidx[k] != j) ++k;
    if (k < e && vals[k] != T(0)) d = T(1) / vals[k];
  }
  invdiag[j] = d;
}

// vals_out[k] = vals_in[src[k]]  (uncompressed input / symmetric expansion)
template <typename T>
__global__ void gather_values_kernel(long long n, const int32_t* __restrict__ src, const T* __restrict__ in,
                                     T* __restrict__ out) {
  for (long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<long long>(gridDim.x) * blockDim.x)
    out[k] = in[src[k]];
}

// ------------------------------------------------------------------------------------------ fused vector passes
// Templated on the scalar type: ConjugateGradient<SparseMatrix<float>> / BiCGSTAB<SparseMatrix<float>> are plain
// instantiations of the same templates in the reference (ConjugateGradient.h:157-160), and so they are here.  Vectors
// and the elementwise arithmetic are in T; dot products accumulate in double (for float operands the products are
// then exact) and are rounded to RealScalar once, in run_epilogue.
template <typename T>
struct VecArgsT {
  long long n;
  T* x;
  T* r;
  T* p;
  T* q;        // Ap (CG) / v (BiCGSTAB)
  T* y;
  T* z;
  T* s;
  T* t;
  T* r0;
  const T* b;
  const T* invdiag;
  double* history;
  RedCtx red;
};

// 128-bit packs: two doubles or four floats.
template <typename T>
struct Pack;
template <>
struct Pack<double> {
  static constexpr int N = 2;
  double v[2];
};
template <>
struct Pack<float> {
  static constexpr int N = 4;
  float v[4];
};

// Streaming 128-bit load: the vector passes touch every element exactly once, so lines are not allocated in L1
// (in the persistent kernel L1 is only what the 4 x 52 KB shared-memory rings leave over).
__device__ __forceinline__ Pack<double> ldp(const double* p, long long ip) {
  Pack<double> v;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
               : "=d"(v.v[0]), "=d"(v.v[1])
               : "l"(reinterpret_cast<const double2*>(p) + ip)
               : "memory");
  return v;
}
__device__ __forceinline__ Pack<float> ldp(const float* p, long long ip) {
  Pack<float> v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.v[0]), "=f"(v.v[1]), "=f"(v.v[2]), "=f"(v.v[3])
               : "l"(reinterpret_cast<const float4*>(p) + ip)
               : "memory");
  return v;
}
__device__ __forceinline__ void stp(double* p, long long ip, const Pack<double>& v) {
  reinterpret_cast<double2*>(p)[ip] = make_double2(v.v[0], v.v[1]);
}
__device__ __forceinline__ void stp(float* p, long long ip, const Pack<float>& v) {
  reinterpret_cast<float4*>(p)[ip] = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]);
}
template <typename T>
__device__ __forceinline__ Pack<T> pack_fill(T c) {
  Pack<T> v;
#pragma unroll
  for (int j = 0; j < Pack<T>::N; ++j) v.v[j] = c;
  return v;
}
// double accumulation of a product of two T values (exact product for float)
__device__ __forceinline__ double dacc(double a, double b, double acc) { return fma_rn(a, b, acc); }
__device__ __forceinline__ double dacc(float a, float b, double acc) {
  return fma_rn(static_cast<double>(a), static_cast<double>(b), acc);
}

// Elements are dealt to threads in 128-bit packs in a fixed grid-stride order; the tail (n mod pack) is taken by
// thread 0 of CTA 0.  fp(ip) handles pack ip, f1(i) a single element.
template <typename T, typename FP, typename F1>
__device__ __forceinline__ void vec_loop(long long n, FP fp, F1 f1) {
  // Two packs per thread and trip: twice the bytes in flight per thread (the order in which a thread visits its
  // elements, hence every reduction, is unchanged).
  constexpr int U = 2;
  constexpr int PN = Pack<T>::N;
  const long long np = n / PN;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long ip = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; ip < np; ip += U * stride) {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (ip + u * stride < np) fp(ip + u * stride);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = np * PN; i < n; ++i) f1(i);
}

// CG start (ConjugateGradient.h:43-67): r = b - A x0 (q holds A x0 when there is a guess, else r = b),
// p = D^-1 r, and the three reductions ||b||^2, ||r||^2, r.p in the same pass.
template <typename T>
__global__ void __launch_bounds__(kVecThreads) cg_init_kernel(const VecArgsT<T> a) {
  pdl_wait();
  __shared__ double scratch[32 * 3];
  const bool guess = a.red.S->use_guess != 0;
  double v[3] = {0.0, 0.0, 0.0};
  auto elem = [&](T b, T q, T d, T& r, T& p) {
    r = guess ? b - q : b;
    p = d * r;
    v[0] = dacc(b, b, v[0]);
    v[1] = dacc(r, r, v[1]);
    v[2] = dacc(r, p, v[2]);
  };
  vec_loop<T>(a.n,
    [&](long long ip) {
      const Pack<T> b = ldp(a.b, ip), d = ldp(a.invdiag, ip);
      Pack<T> q = pack_fill<T>(T(0)), r, p;
      if (guess) q = ldp(a.q, ip);
      else stp(a.x, ip, pack_fill<T>(T(0)));  // solve() starts from x = 0 (IterativeSolverBase.h:402)
#pragma unroll
      for (int j = 0; j < Pack<T>::N; ++j) elem(b.v[j], q.v[j], d.v[j], r.v[j], p.v[j]);
      stp(a.r, ip, r); stp(a.p, ip, p);
    },
    [&](long long i) {
      T r, p;
      elem(a.b[i], guess ? a.q[i] : T(0), a.invdiag[i], r, p);
      if (!guess) a.x[i] = T(0);
      a.r[i] = r; a.p[i] = p;
    });
  pdl_launch_dependents();
  finish_reduction<3, kVecThreads>(a.red, v, scratch, a.history);
}

// CG :75-84 in one pass: r -= alpha Ap; z = D^-1 r (not stored); ||r||^2; r.z.
// The solution update x += alpha p (:74) is DEFERRED to the direction pass, which reads p anyway: one vector read
// less per iteration (12 instead of 13 vector passes of 8N bytes).  Same operations on the same operands, so x is
// bit-identical to the immediate update.
template <typename T>
__device__ __forceinline__ void cg_update_body(const VecArgsT<T>& av, const T alpha, double (&v)[2]) {
  struct { T* __restrict__ r; const T* __restrict__ q; const T* __restrict__ invdiag; } a = {av.r, av.q, av.invdiag};
  auto elem = [&](T& r, T q, T d) {
    r = fma_rn(-alpha, q, r);
    const T z = d * r;
    v[0] = dacc(r, r, v[0]);
    v[1] = dacc(r, z, v[1]);
  };
  vec_loop<T>(av.n,
    [&](long long ip) {
      Pack<T> r = ldp(a.r, ip);
      const Pack<T> q = ldp(a.q, ip), d = ldp(a.invdiag, ip);
#pragma unroll
      for (int j = 0; j < Pack<T>::N; ++j) elem(r.v[j], q.v[j], d.v[j]);
      stp(a.r, ip, r);
    },
    [&](long long i) {
      T r = a.r[i];
      elem(r, a.q[i], a.invdiag[i]);
      a.r[i] = r;
    });
}

template <typename T>
__global__ void __launch_bounds__(kVecThreads) cg_update_kernel(const VecArgsT<T> a) {
  pdl_wait();
  __shared__ double scratch[32 * 2];
  if (gated_out(a.red.S, a.red.gate)) return;
  double v[2] = {0.0, 0.0};
  cg_update_body<T>(a, static_cast<T>(a.red.S->alpha), v);
  pdl_launch_dependents();
  finish_reduction<2, kVecThreads>(a.red, v, scratch, a.history);
}

// CG :74 (deferred) and :81,:86: x += alpha p, then -- unless the loop has just stopped -- p = D^-1 r + beta p
enum { kDirBoth = 0, kDirXOnly = 1, kDirPOnly = 2 };
template <typename T>
__device__ __forceinline__ void cg_direction_body(const VecArgsT<T>& av, const T alpha, const T beta, const int what) {
  struct { T* __restrict__ x; T* __restrict__ p; const T* __restrict__ r; const T* __restrict__ invdiag; } a = {
      av.x, av.p, av.r, av.invdiag};
  if (what == kDirBoth) {
    vec_loop<T>(av.n,
      [&](long long ip) {
        const Pack<T> r = ldp(a.r, ip), d = ldp(a.invdiag, ip);
        Pack<T> p = ldp(a.p, ip), x = ldp(a.x, ip);
#pragma unroll
        for (int j = 0; j < Pack<T>::N; ++j) {
          x.v[j] = fma_rn(alpha, p.v[j], x.v[j]);
          p.v[j] = fma_rn(beta, p.v[j], d.v[j] * r.v[j]);
        }
        stp(a.x, ip, x); stp(a.p, ip, p);
      },
      [&](long long i) {
        const T p = a.p[i];
        a.x[i] = fma_rn(alpha, p, a.x[i]);
        a.p[i] = fma_rn(beta, p, a.invdiag[i] * a.r[i]);
      });
  } else if (what == kDirXOnly) {
    vec_loop<T>(av.n,
      [&](long long ip) {
        const Pack<T> p = ldp(a.p, ip);
        Pack<T> x = ldp(a.x, ip);
#pragma unroll
        for (int j = 0; j < Pack<T>::N; ++j) x.v[j] = fma_rn(alpha, p.v[j], x.v[j]);
        stp(a.x, ip, x);
      },
      [&](long long i) { a.x[i] = fma_rn(alpha, a.p[i], a.x[i]); });
  } else {
    vec_loop<T>(av.n,
      [&](long long ip) {
        const Pack<T> r = ldp(a.r, ip), d = ldp(a.invdiag, ip);
        Pack<T> p = ldp(a.p, ip);
#pragma unroll
        for (int j = 0; j < Pack<T>::N; ++j) p.v[j] = fma_rn(beta, p.v[j], d.v[j] * r.v[j]);
        stp(a.p, ip, p);
      },
      [&](long long i) { a.p[i] = fma_rn(beta, a.p[i], a.invdiag[i] * a.r[i]); });
  }
}

// Runs after every cg_update: applies the pending x update exactly once (n_update counts updates, n_xapplied the
// ones already folded into x; launches that find nothing pending -- gated copies after the stop -- do nothing).
//
// Under programmatic dependent launch (early_x) the CTAs of this kernel become resident while the last CTA of
// cg_update is still folding the partials and exchanging {||r||^2, r.z} with the other ranks (4-5 us per iteration on
// 8 GPUs).  x += alpha p needs nothing from that reduction -- alpha dates from the product before, x and p are not
// touched by cg_update -- so it is done BEFORE griddepcontrol.wait, hidden behind the all-reduce; only p = z + beta p
// waits for beta.  The early part runs only when the loop was live when cg_update started (stop == 0): then cg_update
// is not gated off and this iteration's x update is owed for certain.  Same operation on the same operands: x is
// bit-identical either way.
template <typename T>
__global__ void __launch_bounds__(kVecThreads) cg_direction_kernel(const VecArgsT<T> a, unsigned int* ticket, int early_x) {
  __shared__ int s_early;
  const Scalars* S = a.red.S;
  if (threadIdx.x == 0) s_early = early_x && (__ldcg(&S->stop) == 0);
  __syncthreads();
  const bool early = s_early != 0;
  if (early) cg_direction_body<T>(a, static_cast<T>(__ldcg(&S->alpha)), T(0), kDirXOnly);
  pdl_wait();
  if (__ldcg(&S->n_update) == __ldcg(&S->n_xapplied)) return;
  const bool update_p = __ldcg(&S->stop) == 0;
  const T alpha = static_cast<T>(__ldcg(&S->alpha)), beta = static_cast<T>(__ldcg(&S->beta));
  if (!early) cg_direction_body<T>(a, alpha, beta, update_p ? kDirBoth : kDirXOnly);
  else if (update_p) cg_direction_body<T>(a, alpha, beta, kDirPOnly);
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {  // every CTA has read the counters long ago: safe to retire the pending update
      *ticket = 0;
      a.red.S->n_xapplied = S->n_update;
      timeline_mark(a.red.S, 10);  // interval "update epilogue -> end of the direction pass"
    }
  }
}

// BiCGSTAB start (BiCGSTAB.h:42-46): r = b - A x0 (t holds A x0), r0 = r, ||b||^2, ||r||^2; v = p = 0 (:56)
template <typename T>
__global__ void __launch_bounds__(kVecThreads) bicg_init_kernel(const VecArgsT<T> a) {
  pdl_wait();
  __shared__ double scratch[32 * 2];
  const bool guess = a.red.S->use_guess != 0;
  double v[2] = {0.0, 0.0};
  vec_loop<T>(a.n,
    [&](long long ip) {
      const Pack<T> b = ldp(a.b, ip);
      Pack<T> r = b;
      if (guess) {
        const Pack<T> t = ldp(a.t, ip);
#pragma unroll
        for (int j = 0; j < Pack<T>::N; ++j) r.v[j] = b.v[j] - t.v[j];
      } else {
        stp(a.x, ip, pack_fill<T>(T(0)));
      }
      stp(a.r, ip, r); stp(a.r0, ip, r);
      stp(a.q, ip, pack_fill<T>(T(0))); stp(a.p, ip, pack_fill<T>(T(0)));
#pragma unroll
      for (int j = 0; j < Pack<T>::N; ++j) { v[0] = dacc(b.v[j], b.v[j], v[0]); v[1] = dacc(r.v[j], r.v[j], v[1]); }
    },
    [&](long long i) {
      const T b = a.b[i];
      const T r = guess ? b - a.t[i] : b;
      if (!guess) a.x[i] = T(0);
      a.r[i] = r; a.r0[i] = r; a.q[i] = T(0); a.p[i] = T(0);
      v[0] = dacc(b, b, v[0]); v[1] = dacc(r, r, v[1]);
    });
  pdl_launch_dependents();
  finish_reduction<2, kVecThreads>(a.red, v, scratch, a.history);
}

// BiCGSTAB restart (:75-77): r = b - A x (t holds A x), r0 = r, ||r||^2
template <typename T>
__global__ void __launch_bounds__(kVecThreads) bicg_restart_kernel(const VecArgsT<T> a) {
  pdl_wait();
  __shared__ double scratch[32];
  if (gated_out(a.red.S, a.red.gate)) return;
  double v[1] = {0.0};
  vec_loop<T>(a.n,
    [&](long long ip) {
      const Pack<T> b = ldp(a.b, ip), t = ldp(a.t, ip);
      Pack<T> r;
#pragma unroll
      for (int j = 0; j < Pack<T>::N; ++j) { r.v[j] = b.v[j] - t.v[j]; v[0] = dacc(r.v[j], r.v[j], v[0]); }
      stp(a.r, ip, r); stp(a.r0, ip, r);
    },
    [&](long long i) {
      const T r = a.b[i] - a.t[i];
      a.r[i] = r; a.r0[i] = r;
      v[0] = dacc(r, r, v[0]);
    });
  pdl_launch_dependents();
  finish_reduction<1, kVecThreads>(a.red, v, scratch, a.history);
}

// BiCGSTAB :82-85: beta = (rho/rho_old)(alpha/w); p = r + beta (p - w v); y = D^-1 p
template <typename T>
__global__ void __launch_bounds__(kVecThreads) bicg_p_kernel(const VecArgsT<T> a) {
  pdl_wait();
  if (gated_out(a.red.S, a.red.gate)) return;
  const Scalars* S = a.red.S;
  // every factor is already a RealScalar value; the two quotients and the product are T operations (:82)
  const T beta = (static_cast<T>(S->rho) / static_cast<T>(S->rho_old)) * (static_cast<T>(S->alpha) / static_cast<T>(S->w));
  const T w = static_cast<T>(S->w);
  vec_loop<T>(a.n,
    [&](long long ip) {
      const Pack<T> r = ldp(a.r, ip), vv = ldp(a.q, ip), d = ldp(a.invdiag, ip);
      Pack<T> p = ldp(a.p, ip), y;
#pragma unroll
      for (int j = 0; j < Pack<T>::N; ++j) {
        p.v[j] = fma_rn(beta, fma_rn(-w, vv.v[j], p.v[j]), r.v[j]);
        y.v[j] = d.v[j] * p.v[j];
      }
      stp(a.p, ip, p);
      stp(a.y, ip, y);
    },
    [&](long long i) {
      const T p = fma_rn(beta, fma_rn(-w, a.q[i], a.p[i]), a.r[i]);
      a.p[i] = p; a.y[i] = a.invdiag[i] * p;
    });
}

// BiCGSTAB :90-92: s = r - alpha v; z = D^-1 s
template <typename T>
__global__ void __launch_bounds__(kVecThreads) bicg_s_kernel(const VecArgsT<T> a) {
  pdl_wait();
  if (gated_out(a.red.S, a.red.gate)) return;
  const T alpha = static_cast<T>(a.red.S->alpha);
  vec_loop<T>(a.n,
    [&](long long ip) {
      const Pack<T> r = ldp(a.r, ip), vv = ldp(a.q, ip), d = ldp(a.invdiag, ip);
      Pack<T> s, z;
#pragma unroll
      for (int j = 0; j < Pack<T>::N; ++j) {
        s.v[j] = fma_rn(-alpha, vv.v[j], r.v[j]);
        z.v[j] = d.v[j] * s.v[j];
      }
      stp(a.s, ip, s);
      stp(a.z, ip, z);
    },
    [&](long long i) {
      const T s = fma_rn(-alpha, a.q[i], a.r[i]);
      a.s[i] = s; a.z[i] = a.invdiag[i] * s;
    });
}

// BiCGSTAB :100-101 and the next loop head :67,:71: x += alpha y + w z; r = s - w t; ||r||^2; r0.r
template <typename T>
__global__ void __launch_bounds__(kVecThreads) bicg_update_kernel(const VecArgsT<T> a) {
  pdl_wait();
  __shared__ double scratch[32 * 2];
  if (gated_out(a.red.S, a.red.gate)) return;
  const T alpha = static_cast<T>(a.red.S->alpha), w = static_cast<T>(a.red.S->w);
  double v[2] = {0.0, 0.0};
  auto elem = [&](T& x, T y, T z, T s, T t, T r0, T& r) {
    x = x + fma_rn(w, z, alpha * y);
    r = fma_rn(-w, t, s);
    v[0] = dacc(r, r, v[0]);
    v[1] = dacc(r0, r, v[1]);
  };
  vec_loop<T>(a.n,
    [&](long long ip) {
      Pack<T> x = ldp(a.x, ip), r;
      const Pack<T> y = ldp(a.y, ip), z = ldp(a.z, ip), s = ldp(a.s, ip), t = ldp(a.t, ip), r0 = ldp(a.r0, ip);
#pragma unroll
      for (int j = 0; j < Pack<T>::N; ++j) elem(x.v[j], y.v[j], z.v[j], s.v[j], t.v[j], r0.v[j], r.v[j]);
      stp(a.x, ip, x); stp(a.r, ip, r);
    },
    [&](long long i) {
      T x = a.x[i], r;
      elem(x, a.y[i], a.z[i], a.s[i], a.t[i], a.r0[i], r);
      a.x[i] = x; a.r[i] = r;
    });
  pdl_launch_dependents();
  finish_reduction<2, kVecThreads>(a.red, v, scratch, a.history);
}

// x = 0 when ||b|| == 0 (ConjugateGradient.h:48, BiCGSTAB.h:49); x = NaN when CG ran into a non-finite residual norm
// with at least the iterations left that the reference needs to spread it over x (see kEpiCgInit); otherwise nothing.
template <typename T>
__global__ void __launch_bounds__(kVecThreads) finalize_kernel(const VecArgsT<T> a) {
  pdl_wait();
  const bool zero = a.red.S->rhs_zero != 0, nan = a.red.S->numerical_issue == 2;
  if (!zero && !nan) return;
  const T fill = zero ? T(0) : static_cast<T>(__longlong_as_double(0x7ff8000000000000ll));
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    a.x[i] = fill;
}

// Sets the WHILE condition from the control state (used after the init phase and by chunk boundaries).
__global__ void set_condition_kernel(const Scalars* S, unsigned long long handle) {
  cudaGraphSetConditional(static_cast<cudaGraphConditionalHandle>(handle), S->stop ? 0u : 1u);
}


// ------------------------------------------------------------------------------------- persistent CG (one launch)
// The whole loop of ConjugateGradient.h:69-88 in ONE cooperative kernel: the three passes of an iteration are phases
// separated by grid-wide barriers instead of kernel boundaries, and the reductions ride on the barriers (the CTA
// that arrives last folds the partials, all-reduces over ranks, runs the scalar epilogue, then releases everybody).
// Same device functions and the same arithmetic as the three-kernel pipeline; the vector phases run on this kernel's
// grid, so the grouping of the partial sums (hence the last bits of the dot products) matches the graph modes only
// while both grids cover the vector in one sweep (n/2 <= grid * 256); beyond that the two are equally deterministic
// but not bit-identical to each other.  What disappears is launch latency, which dominates when an iteration is
// tens of microseconds (strong scaling over 8 GPUs, L2-sized problems).
template <typename T>
struct CgPersistArgs {
  SpmvArgs<T> sp;
  VecArgsT<T> ve;
  unsigned int* bar_count;
  unsigned int* bar_gen;
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Grid barrier; with NV > 0 also a deterministic reduction of v over CTAs and ranks followed by the epilogue.
template <int NV, int THREADS>
__device__ __forceinline__ void grid_sync(const RedCtx& ctx, double* v, double* scratch, double* history,
                                          unsigned* bar_count, unsigned* bar_gen, unsigned& gen) {
  __shared__ int s_last_g;
  if (NV > 0) {
    double t[NV > 0 ? NV : 1];
#pragma unroll
    for (int j = 0; j < NV; ++j) t[j] = v[j];
    block_reduce<(NV > 0 ? NV : 1), THREADS>(t, scratch);
    if (threadIdx.x == 0)
#pragma unroll
      for (int j = 0; j < NV; ++j) ctx.partials[blockIdx.x * 4 + j] = t[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned ticket = atomicAdd(bar_count, 1u);
    s_last_g = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last_g) {
    double r[4] = {0, 0, 0, 0};
    if (NV > 0) {
      __threadfence();
      double t[NV > 0 ? NV : 1];
#pragma unroll
      for (int j = 0; j < NV; ++j) t[j] = 0.0;
      for (unsigned b = threadIdx.x; b < gridDim.x; b += THREADS)
#pragma unroll
        for (int j = 0; j < NV; ++j) t[j] += __ldcg(ctx.partials + b * 4 + j);
      block_reduce<(NV > 0 ? NV : 1), THREADS>(t, scratch);
#pragma unroll
      for (int j = 0; j < NV; ++j) r[j] = t[j];
      allreduce_ranks(ctx.comm, ctx.S, r, NV);
    }
    if (threadIdx.x == 0) {
      *bar_count = 0;
      if (NV > 0) {
        if (ctx.bump_halo) ctx.S->halo_seq++;
        timeline_mark(ctx.S, ctx.epilogue);
        run_epilogue(ctx, r, history);
      }
      if (ctx.S->comm_error) ctx.S->stop = 1;
      __threadfence();
      st_release_gpu(bar_gen, gen + 1);
    }
  } else if (threadIdx.x == 0) {
    unsigned spins = 0;
    const unsigned long long tw = globaltimer_ns();
    while (ld_acquire_gpu(bar_gen) != gen + 1) {  // back off: 591 pollers share one L2 line
      __nanosleep(40);
      if (spin_expired(ctx.S, tw, spins)) break;
    }
    __threadfence();
  }
  __syncthreads();
  ++gen;
}

template <typename T>
__global__ void __launch_bounds__(kSpmvThreads, 1024 / kSpmvThreads) cg_persistent_kernel(const CgPersistArgs<T> a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[8];
  __shared__ double red_scratch[32 * 2];
  __shared__ T long_scratch[32];
  Scalars* S = a.sp.red.S;
  SpmvCta<T> cx;
  spmv_cta_init(a.sp, cx, smem, full_bar, long_scratch);
  unsigned gen = __ldcg(a.bar_gen);  // stable: only barriers of THIS kernel advance it, and nobody has arrived yet
  // every CTA must have read `gen` before anyone can release the first barrier: guaranteed, because a release needs
  // all CTAs to arrive, and each arrives after this read.
  RedCtx red_pap = a.sp.red;
  red_pap.epilogue = kEpiCgPAp;
  red_pap.bump_halo = a.sp.halo.enabled;
  RedCtx red_upd = a.ve.red;
  red_upd.epilogue = kEpiCgUpdate;
  red_upd.bump_halo = 0;
  RedCtx red_none = a.ve.red;
  red_none.epilogue = kEpiNone;

  bool first = true, pending = false;
  if (!__ldcg(&S->stop)) {
    spmv_prefetch(a.sp, cx);
    pending = true;
  }
  // S changes only inside barrier epilogues, so after every barrier all CTAs read the same control state
  while (!__ldcg(&S->stop)) {
    if (!first) {
      cg_direction_body<T>(a.ve, static_cast<T>(__ldcg(&S->alpha)), static_cast<T>(__ldcg(&S->beta)), kDirBoth);
      grid_sync<0, kSpmvThreads>(red_none, nullptr, red_scratch, a.ve.history, a.bar_count, a.bar_gen, gen);
    }
    first = false;
    if (a.sp.halo.enabled) halo_push<T>(a.sp.halo, a.sp.red.comm, S, a.sp.x);
    double d0 = 0.0, d1 = 0.0;
    spmv_tiles<T, 1, false>(a.sp, cx, d0, d1);
    spmv_prefetch(a.sp, cx);  // the next product's first tiles fly during the two vector phases
    {
      double v[1] = {d0};
      grid_sync<1, kSpmvThreads>(red_pap, v, red_scratch, a.ve.history, a.bar_count, a.bar_gen, gen);
    }
    {
      double v[2] = {0.0, 0.0};
      cg_update_body<T>(a.ve, static_cast<T>(__ldcg(&S->alpha)), v);
      grid_sync<2, kSpmvThreads>(red_upd, v, red_scratch, a.ve.history, a.bar_count, a.bar_gen, gen);
    }
  }
  if (!first) {  // the last iteration's x += alpha p is still owed (the loop stopped before its direction pass)
    cg_direction_body<T>(a.ve, static_cast<T>(__ldcg(&S->alpha)), T(0), kDirXOnly);
    if (blockIdx.x == 0 && threadIdx.x == 0) S->n_xapplied = S->n_update;
  }
  if (pending) {  // tiles prefetched for a product that will not happen: wait until the copies have landed
    for (int k = 0; k < a.sp.stages; ++k) {
      const int t = blockIdx.x + k * gridDim.x;
      if (t < a.sp.ntiles) {
        const unsigned q = cx.seq + k;
        mbar_wait(&full_bar[q % a.sp.stages], (q / a.sp.stages) & 1);
      }
    }
  }
}

}  // namespace b200s

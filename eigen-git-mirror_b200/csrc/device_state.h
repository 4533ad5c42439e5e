// device_state.h -- structures shared by host code and kernels.
#pragma once
#include <cstdint>

#include "plan.h"

namespace b200s {

constexpr int kVecThreads = 256;      // CTA size of the fused vector kernels
constexpr int kMaxGrid = 2048;        // upper bound of any reduction grid (partials per reduction slot)
constexpr int kMaxWorld = 8;          // GPUs of one NVSwitch box
constexpr int kHistoryCap = 1 << 16;  // residual-history ring (doubles)
constexpr int kSolveInputBytes = 32;  // leading bytes of Scalars the host writes before every solve

// Solver state that lives in device memory for the whole solve.  Only "last blocks" (the CTA that takes the final
// ticket of a reduction) and single-thread control kernels write it, so every kernel of an iteration reads a
// consistent snapshot and takes the same branches as the host loop of the reference would.
struct Scalars {
  // inputs of the solve (written by the host before the graph is launched)
  double tol;
  long long max_iters;
  int use_guess;
  int comm_error;  // a bounded spin expired (dead or diverged peer): the loop was stopped, results are invalid
  unsigned long long comm_timeout_ns;  // bound of every cross-rank / grid-wide spin (0 = default 20 s)
  // shared by CG and BiCGSTAB
  double bb;       // ||b||^2                       ConjugateGradient.h:45 / BiCGSTAB.h:46
  double thr;      // CG: max(tol^2 bb, DBL_MIN) :53-54 ; BiCGSTAB: tol^2 bb :62
  double rr;       // ||r||^2 of the current residual
  long long iter;  // completed iterations (the reference's `i`)
  int stop;        // loop finished (converged, exhausted, trivial rhs, numerical issue)
  int converged;
  int rhs_zero;    // ||b|| == 0 : x is zeroed at the end
  int numerical_issue;
  // CG
  double rz, abs_new, abs_old, pAp, alpha, beta;
  long long n_update, n_xapplied;  // deferred x += alpha p bookkeeping (see cg_direction_kernel)
  // BiCGSTAB
  double rho, rho_old, w, r0_sqnorm, r0v, ts, tt, rho_next, eps2;
  long long restarts;
  int restart;     // this iteration starts with the re-orthogonalisation branch (BiCGSTAB.h:72-81)
  int stop_all;    // multi-column solves: every column's `stop` is set (kept in the first column's block)
  // bookkeeping
  long long spmv_count;
  long long hist_len;
  unsigned int red_seq;   // cross-rank reduction sequence number (multi-GPU mailboxes)
  unsigned int halo_seq;  // halo-exchange sequence number
  // device-side timeline (globaltimer ns), accumulated per solve: where an iteration's time goes
  unsigned long long t_last;          // time of the previous reduction epilogue
  unsigned long long t_phase[12];     // [epilogue kind] time since the previous epilogue (kernel + launch gap)
  unsigned long long n_phase[12];
  unsigned long long t_allreduce;     // spent inside the cross-rank all-reduce (stores, fence, waiting for peers)
  unsigned long long t_halo_wait;     // spent by CTA 0 waiting for ghost entries
  unsigned long long t_first, t_end;
};

// What the last block of a reduction does with the reduced values.
enum Epilogue : int {
  kEpiNone = 0,
  kEpiSpmvOnly,      // plain y = A x (no scalar logic)
  kEpiCgInit,        // bb, rr, rz  -> thresholds, early outs, abs_new
  kEpiCgPAp,         // p.Ap        -> alpha
  kEpiCgUpdate,      // rr, rz      -> convergence test, beta, iter++
  kEpiBiInit,        // bb, rr      -> thresholds, rho = r0_sqnorm = rr
  kEpiBiR0V,         // r0.v        -> alpha
  kEpiBiTsTt,        // t.s, t.t    -> w
  kEpiBiUpdate,      // rr, r0.r    -> iter++, loop test, restart test
  kEpiBiRestart,     // rr          -> rho = r0_sqnorm = rr, restarts++
};

// Kernel gating on the device-resident control state.
enum Gate : int {
  kGateNone = 0,
  kGateLoop,     // skip when S->stop
  kGateRestart,  // run only when !S->stop && S->restart
  kGateGuess,    // run only when S->use_guess (initial residual needs A*x0)
};

// Peer-memory window of one rank (multi-GPU); all pointers are device pointers valid on THIS device: `self` ones
// in local HBM, `peer` ones mapped through CUDA IPC over NVLink.
struct CommDev {
  int world;
  int rank;
  // mailboxes for the scalar all-reduce: box[parity][src_rank][4 doubles], flags[parity][src_rank]
  double* box_self;
  unsigned int* flag_self;
  double* box_peer[kMaxWorld];
  unsigned int* flag_peer[kMaxWorld];
  // halo arrival flags: halo_flag_self[src_rank] = sequence number of the last completed push from src_rank
  unsigned int* halo_flag_self;
  unsigned int* halo_flag_peer[kMaxWorld];
};

struct RedCtx {
  double* partials;        // [kMaxGrid * 4]
  unsigned int* counter;   // ticket counter, self-resetting
  Scalars* S;
  int epilogue;
  int gate;
  unsigned long long cond_handle;  // cudaGraphConditionalHandle or 0
  int set_cond;                    // last block updates the WHILE condition after the epilogue
  int bump_halo;                   // the kernel carried a halo exchange (multi-GPU SpMV)
  int f32;                         // RealScalar = float: the scalar epilogue rounds every operation to float
  CommDev comm;
};

}  // namespace b200s

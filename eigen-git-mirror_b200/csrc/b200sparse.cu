// b200sparse.cu -- C ABI (include/b200sparse.h) over the sm_100a kernels in kernels.cuh.
//
// Host-side structure of one handle (= one Eigen solver object, IterativeSolverBase.h:142-440):
//   analyze_pattern : plan (plan.cpp) -> device pattern, tiles, peer window (multi-GPU), launch geometry
//   factorize       : values -> device, Jacobi inverse diagonal (BasicPreconditioners.h:64-79)
//   *_solve         : b/x0 -> device, ONE CUDA-graph launch whose WHILE node iterates until the device-resident
//                     control state says stop, x -> host.  No host synchronisation inside the iteration.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/b200sparse.h"
#include "kernels.cuh"
#include "kernels_multi.cuh"
#include "kernels_krylov.cuh"
#include "kernels_tri.cuh"
#include "factors.h"

using namespace b200s;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  bool owned = true;  // false: a slice of the handle's vector window
  template <typename T>
  T* as() const { return static_cast<T*>(p); }
};

struct GraphSet {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaGraph_t body_graph = nullptr;   // chunked mode: the unrolled body
  cudaGraphExec_t body_exec = nullptr;
  int init_kernels = 0, body_kernels = 0, tail_kernels = 0, body_unroll = 1;
  bool built = false;
  const void* tag = nullptr;  // address the graph's kernels were built around (rebuilt when it moves)
};

// device copy of an incomplete factorization (csrc/factors.h): two level-scheduled triangular stages between
// a gather/scale on the way in and one on the way out
struct DevTriStage {
  DevBuf rowptr, colidx, vals, diag, level_ptr, level_rows;
  bool unit = true, fused = false;
  int levels = 0;
  std::vector<b200s::TriStage::Launch> launches;
};
struct DevFactors {
  bool set = false;
  int kind = 0;
  int64_t n = 0;
  DevTriStage st[2];
  DevBuf pre_gather, post_gather, pre_scale, post_scale, x;
  int64_t applies = 0;
};

}  // namespace

// the opaque factors object of the C ABI is the host-side analysis result
struct b200s_factors {
  b200s::Factors f;
};

struct b200s_handle {
  b200s_config cfg{};
  int device = 0, sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  Plan plan;
  bool analyzed = false, factorized = false;
  int scalar_bytes = 0;  // 8 = f64, 4 = f32
  int precond = B200S_PRECOND_JACOBI;
  int precond_factorize = B200S_PRECOND_JACOBI;  // what factorize() built (set_preconditioner(NULL) returns to it)
  DevFactors fac;
  int loop_mode = B200S_LOOP_WHILE_GRAPH;  // resolved per problem in analyze_pattern when AUTO was requested
  bool loop_auto = false;
  int spmv_impl = B200S_SPMV_STAGED;
  int direct_lg = 0;
  // device pattern / values
  DevBuf rowptr, colidx, vals, src, tiles, invdiag, send_rows;
  // vectors (doubles unless noted); ext = [owned | ghost]
  DevBuf window;  // peer-visible: 4 extended slots + mailboxes + flags
  size_t slot_bytes = 0, slots_off = 0, box_off = 0, flag_off = 0, halo_flag_off = 0, window_bytes = 0;
  int l2_persist = 0;  // the front of the window (r, Ap, D^-1, x, p) is marked persisting in L2
  std::vector<void*> peer_window;  // [world], IPC-mapped (self: window.p)
  std::vector<int64_t> all_rows, all_ghosts;
  DevBuf b, r, q, r0, s, t, yout;
  DevBuf scalars, partials, counter, halo_counter, history, gridbar;
  int persist_grid = 0, persist_grid_f32 = 0;
  Scalars* hS = nullptr;  // pinned mirror
  double comm_timeout_ms = 20000.0;  // bound of every device-side wait on a peer (B200S_COMM_TIMEOUT_MS)
  // launch geometry
  int spmv_grid = 0, spmv_stages = 0, spmv_stages_f32 = 0, spmv_smem = 0, vec_grid = 0;
  int spmv_grid_f32 = 0, spmv_smem_f32 = 0;  // float tiles are smaller: more CTAs fit per SM
  int evict_first = 0;
  int pdl = 0;           // programmatic dependent launch between the solver kernels (B200S_PDL, default on)
  int early_x = 0;       // cg_direction applies x += alpha p before its grid dependency resolves (see kernels.cuh)
  int body_unroll = 1;   // iterations per WHILE-body (amortises the loop-back, keeps PDL edges inside the body)
  GraphSet cg, bicg;
  // multi-column CG (kernels_multi.cuh): interleaved [rows][K] vectors, one control block per column
  GraphSet cg_multi[4];  // K = 2, 4, 8 -> index log2(K)
  DevBuf mx, mr, mp, mq, mb, mS, mpartials, mstage;
  Scalars* hSm = nullptr;  // pinned mirror of the K control blocks
  size_t device_bytes = 0;
  // stats of the last call
  double last_solve_ms = 0, last_h2d_ms = 0, last_d2h_ms = 0;
  int64_t last_launches = 0, last_iterations = 0, last_spmv = 0, last_restarts = 0;
  int last_nonfinite = 0, last_comm_error = 0;
};

namespace {

#define CK(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) {                                                                            \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"; \
      return B200S_ERR_CUDA;                                                                            \
    }                                                                                                   \
  } while (0)

int fail(b200s_handle* h, int code, const std::string& msg) {
  h->err = msg;
  return code;
}

void free_factors(b200s_handle* h);  // precond.inc

int dev_alloc(b200s_handle* h, DevBuf& buf, size_t bytes, bool zero = false) {
  bytes = ((bytes + 255) & ~size_t(255)) + 256;  // slack: bulk copies read up to 16 B past the logical end
  if (buf.p && buf.bytes >= bytes) {
    if (zero) CK(cudaMemsetAsync(buf.p, 0, buf.bytes, h->stream));
    return 0;
  }
  if (buf.p && !buf.owned) return fail(h, B200S_ERR_INVALID, "internal: window slice too small");
  if (buf.p) {
    cudaFree(buf.p);
    h->device_bytes -= buf.bytes;
    buf.p = nullptr;
    buf.bytes = 0;
  }
  cudaError_t e = cudaMalloc(&buf.p, bytes);
  if (e != cudaSuccess) {
    buf.p = nullptr;
    return fail(h, B200S_ERR_ALLOC, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
  }
  buf.bytes = bytes;
  h->device_bytes += bytes;
  if (zero) CK(cudaMemsetAsync(buf.p, 0, bytes, h->stream));
  return 0;
}

void dev_free(b200s_handle* h, DevBuf& buf) {
  if (buf.p && buf.owned) {
    cudaFree(buf.p);
    h->device_bytes -= buf.bytes;
  }
  buf.p = nullptr;
  buf.bytes = 0;
  buf.owned = true;
}

void destroy_graphs(GraphSet& g) {
  if (g.exec) cudaGraphExecDestroy(g.exec);
  if (g.graph) cudaGraphDestroy(g.graph);
  if (g.body_exec) cudaGraphExecDestroy(g.body_exec);
  if (g.body_graph) cudaGraphDestroy(g.body_graph);
  g = GraphSet();
}

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}

// ---- window layout (identical arithmetic on every rank, for every rank) ----
// ONE allocation per handle holds every vector of the solvers:
//   [ r | q | invdiag | slot X | slot P | slot Z | slot Spmv | mailboxes | flags | halo flags | b | r0 | s | t | yout ]
// The four slots are extended vectors [owned | ghost] that peers store into (the allocation is exported over CUDA
// IPC); the others are private.  The order puts the five vectors a CG iteration touches (r, Ap, D^-1, x, p) at the
// front, contiguous, so that ONE L2 access-policy window can mark exactly them as persisting (see l2_persist).
size_t ext_slot_bytes(int64_t rows, int64_t ghosts) {
  return ((static_cast<size_t>(rows + ghosts) * 8 + 255) & ~size_t(255)) + 256;
}
size_t priv_vec_bytes(int64_t rows) { return ((static_cast<size_t>(rows) * 8 + 255) & ~size_t(255)) + 256; }
enum { kSlotX = 0, kSlotP = 1, kSlotZ = 2, kSlotSpmv = 3, kNumSlots = 4 };
enum { kFrontVecs = 3 /* r, q, invdiag */, kBackVecs = 5 /* b, r0, s, t, yout */ };
size_t slots_off(int64_t rows) { return kFrontVecs * priv_vec_bytes(rows); }

template <typename T>
T* slot_ptr(const b200s_handle* h, int slot) {
  return reinterpret_cast<T*>(static_cast<char*>(h->window.p) + h->slots_off + h->slot_bytes * slot);
}

CommDev make_comm(const b200s_handle* h) {
  CommDev c{};
  c.world = h->plan.world;
  c.rank = h->plan.rank;
  char* self = static_cast<char*>(h->window.p);
  c.box_self = reinterpret_cast<double*>(self + h->box_off);
  c.flag_self = reinterpret_cast<unsigned*>(self + h->flag_off);
  c.halo_flag_self = reinterpret_cast<unsigned*>(self + h->halo_flag_off);
  for (int q = 0; q < c.world && q < kMaxWorld; ++q) {
    // offsets inside a peer's window depend on that peer's slot size
    size_t sb = ext_slot_bytes(h->all_rows[q], h->all_ghosts[q]);
    char* base = static_cast<char*>(h->peer_window[q]);
    size_t box_off = slots_off(h->all_rows[q]) + sb * kNumSlots;
    size_t flag_off = box_off + sizeof(unsigned long long) * 2 * kMaxWorld * kBoxWords;
    size_t halo_off = flag_off + sizeof(unsigned) * 2 * kMaxWorld;
    c.box_peer[q] = reinterpret_cast<double*>(base + box_off);
    c.flag_peer[q] = reinterpret_cast<unsigned*>(base + flag_off);
    c.halo_flag_peer[q] = reinterpret_cast<unsigned*>(base + halo_off);
  }
  return c;
}

RedCtx make_red(const b200s_handle* h, int epilogue, int gate, bool set_cond, cudaGraphConditionalHandle cond) {
  RedCtx r{};
  r.partials = h->partials.as<double>();
  r.counter = h->counter.as<unsigned>();
  r.S = h->scalars.as<Scalars>();
  r.epilogue = epilogue;
  r.gate = gate;
  r.cond_handle = static_cast<unsigned long long>(cond);
  r.set_cond = set_cond ? 1 : 0;
  r.comm = make_comm(h);
  r.f32 = (h->scalar_bytes == 4) ? 1 : 0;
  return r;
}

unsigned recv_mask(const b200s_handle* h) {
  unsigned m = 0;
  for (int q = 0; q < h->plan.world; ++q)
    if (h->plan.recv_counts[q] > 0) m |= 1u << q;
  return m;
}

// ---- kernel launch helpers (all on h->stream; counted) ----
// Solver kernels are launched with programmatic stream serialization (PDL): a kernel may become resident and run its
// data-independent prologue (mbarrier setup, the first tile copies of the SpMV) while its predecessor drains; every
// kernel executes griddepcontrol.wait before it touches anything the predecessor wrote.
template <typename... KArgs, typename... Args>
cudaError_t launch_k(b200s_handle* h, void (*kernel)(KArgs...), int grid, int block, size_t smem, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = h->pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// `halo_slot` >= 0: x_ext is that extended slot of the peer-visible window and its boundary entries are pushed to the
// neighbours at the head of the kernel (multi-GPU only).
template <typename T>
int make_spmv_args(b200s_handle* h, SpmvArgs<T>& a, const T* x_ext, T* y, const T* w, int epilogue, int gate,
                   bool set_cond, cudaGraphConditionalHandle cond, int halo_slot);
template <typename T>
int launch_spmv_args(b200s_handle* h, const SpmvArgs<T>& a, int ndot);

template <typename T>
int launch_spmv(b200s_handle* h, const T* x_ext, T* y, const T* w, int ndot, int epilogue, int gate, bool set_cond,
                cudaGraphConditionalHandle cond, int halo_slot) {
  SpmvArgs<T> a{};
  int rc0 = make_spmv_args<T>(h, a, x_ext, y, w, epilogue, gate, set_cond, cond, halo_slot);
  if (rc0) return rc0;
  return launch_spmv_args<T>(h, a, ndot);
}

template <typename T>
int make_spmv_args(b200s_handle* h, SpmvArgs<T>& a, const T* x_ext, T* y, const T* w, int epilogue, int gate,
                   bool set_cond, cudaGraphConditionalHandle cond, int halo_slot) {
  a.tiles = h->tiles.as<Tile>();
  a.ntiles = static_cast<int>(h->plan.tiles.size());
  a.first_boundary_tile = a.ntiles - h->plan.n_boundary_tiles;
  a.rowptr = h->rowptr.as<int32_t>();
  a.colidx = h->colidx.as<int32_t>();
  a.vals = h->vals.as<T>();
  a.x = x_ext;
  a.y = y;
  a.w = w;
  a.stages = sizeof(T) == 4 ? h->spmv_stages_f32 : h->spmv_stages;
  a.cap_nnz = h->plan.tile_nnz;
  a.cap_rows = h->plan.tile_rows_cap;
  a.tail_blk = (sizeof(T) == 4) ? 8 : 0;
  a.evict_first = h->evict_first;
  a.recv_mask = recv_mask(h);
  a.history = h->history.as<double>();
  if (h->plan.world > 1) {
    if (halo_slot < 0) return fail(h, B200S_ERR_INVALID, "internal: multi-GPU SpMV needs a window slot");
    if (epilogue == kEpiNone) epilogue = kEpiSpmvOnly;  // the final rendezvous retires the halo sequence number
    a.halo.enabled = 1;
    {
      const int grid = (h->spmv_impl == B200S_SPMV_DIRECT) ? h->sm_count * 8
                                                           : (sizeof(T) == 4 ? h->spmv_grid_f32 : h->spmv_grid);
      const int64_t total = static_cast<int64_t>(h->plan.send_rows.size());
      a.halo.npush = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(grid, (total + 2047) / 2048)));
    }
    a.halo.send_rows = h->send_rows.as<int32_t>();
    a.halo.counter = h->halo_counter.as<unsigned>();
    for (int q = 0; q < h->plan.world; ++q) {
      a.halo.send_offsets[q] = h->plan.send_offsets[q];
      a.halo.send_counts[q] = h->plan.send_counts[q];
      size_t sb = ext_slot_bytes(h->all_rows[q], h->all_ghosts[q]);
      char* base = static_cast<char*>(h->peer_window[q]) + slots_off(h->all_rows[q]) + sb * halo_slot;
      a.halo.dst[q] = reinterpret_cast<T*>(base) + h->all_rows[q] + h->plan.send_slot0[q];
    }
  }
  a.red = make_red(h, epilogue, gate, set_cond, cond);
  a.red.bump_halo = a.halo.enabled;
  return 0;
}

template <typename T>
int launch_spmv_args(b200s_handle* h, const SpmvArgs<T>& a, int ndot) {
  if (h->spmv_impl == B200S_SPMV_DIRECT) {
    const int rows = static_cast<int>(h->plan.rows);
    const int grid = h->sm_count * 8;
#define B200S_DIRECT(LG)                                                                                   \
  case LG:                                                                                                 \
    if (ndot == 0) launch_k(h, spmv_direct_kernel<T, LG, 0>, grid, kSpmvThreads, 0, a, rows);            \
    else if (ndot == 1) launch_k(h, spmv_direct_kernel<T, LG, 1>, grid, kSpmvThreads, 0, a, rows);       \
    else launch_k(h, spmv_direct_kernel<T, LG, 2>, grid, kSpmvThreads, 0, a, rows);                      \
    break;
    switch (h->direct_lg) {
      B200S_DIRECT(0) B200S_DIRECT(1) B200S_DIRECT(2) B200S_DIRECT(3) B200S_DIRECT(4)
      default:
        if (ndot == 0) launch_k(h, spmv_direct_kernel<T, 5, 0>, grid, kSpmvThreads, 0, a, rows);
        else if (ndot == 1) launch_k(h, spmv_direct_kernel<T, 5, 1>, grid, kSpmvThreads, 0, a, rows);
        else launch_k(h, spmv_direct_kernel<T, 5, 2>, grid, kSpmvThreads, 0, a, rows);
    }
#undef B200S_DIRECT
  } else {
    const int grid = sizeof(T) == 4 ? h->spmv_grid_f32 : h->spmv_grid;
    const int smem = sizeof(T) == 4 ? h->spmv_smem_f32 : h->spmv_smem;
    if (ndot == 0) launch_k(h, spmv_staged_kernel<T, 0>, grid, kSpmvThreads, smem, a);
    else if (ndot == 1) launch_k(h, spmv_staged_kernel<T, 1>, grid, kSpmvThreads, smem, a);
    else launch_k(h, spmv_staged_kernel<T, 2>, grid, kSpmvThreads, smem, a);
  }
  CK(cudaGetLastError());
  h->last_launches++;
  return 0;
}

template <typename T>
VecArgsT<T> make_vec(b200s_handle* h, int epilogue, int gate, bool set_cond, cudaGraphConditionalHandle cond) {
  VecArgsT<T> a{};
  a.n = h->plan.rows;
  a.x = slot_ptr<T>(h, kSlotX);
  a.p = slot_ptr<T>(h, kSlotP);  // CG: p ; BiCGSTAB: y lives in kSlotP, p in h->yout (not exchanged)
  a.z = slot_ptr<T>(h, kSlotZ);
  a.r = h->r.as<T>();
  a.q = h->q.as<T>();
  a.r0 = h->r0.as<T>();
  a.s = h->s.as<T>();
  a.t = h->t.as<T>();
  a.y = nullptr;
  a.b = h->b.as<T>();
  a.invdiag = h->invdiag.as<T>();
  a.history = h->history.as<double>();
  a.red = make_red(h, epilogue, gate, set_cond, cond);
  return a;
}

#define LAUNCH_VEC(kernel, args)                                     \
  do {                                                               \
    launch_k(h, kernel, h->vec_grid, kVecThreads, 0, args);          \
    CK(cudaGetLastError());                                          \
    h->last_launches++;                                              \
  } while (0)

// ---- the iteration bodies (enqueued either under stream capture or directly) ----
template <typename T>
int enqueue_cg_init(b200s_handle* h, bool set_cond, cudaGraphConditionalHandle cond) {
  int rc;
  if ((rc = launch_spmv<T>(h, slot_ptr<T>(h, kSlotX), h->q.as<T>(), nullptr, 0, kEpiNone, kGateGuess, false, 0, kSlotX)))
    return rc;
  VecArgsT<T> a = make_vec<T>(h, kEpiCgInit, kGateNone, set_cond, cond);
  LAUNCH_VEC(cg_init_kernel<T>, a);
  return 0;
}

template <typename T>
int enqueue_cg_body(b200s_handle* h, bool set_cond, cudaGraphConditionalHandle cond) {
  int rc;
  if ((rc = launch_spmv<T>(h, slot_ptr<T>(h, kSlotP), h->q.as<T>(), nullptr, 1, kEpiCgPAp, kGateLoop, false, 0, kSlotP)))
    return rc;
  VecArgsT<T> u = make_vec<T>(h, kEpiCgUpdate, kGateLoop, set_cond, cond);
  LAUNCH_VEC(cg_update_kernel<T>, u);
  VecArgsT<T> d = make_vec<T>(h, kEpiNone, kGateNone, false, 0);
  CK(launch_k(h, cg_direction_kernel<T>, h->vec_grid, kVecThreads, 0, d, h->gridbar.as<unsigned>() + 64,
              (h->pdl && h->early_x) ? 1 : 0));
  h->last_launches++;
  return 0;
}

// The whole CG loop as one cooperative kernel (B200S_LOOP_PERSISTENT).
template <typename T>
int launch_cg_persistent(b200s_handle* h) {
  CgPersistArgs<T> a{};
  int rc = make_spmv_args<T>(h, a.sp, slot_ptr<T>(h, kSlotP), h->q.as<T>(), nullptr, kEpiCgPAp, kGateNone, false, 0,
                             kSlotP);
  if (rc) return rc;
  const int grid = sizeof(T) == 4 ? h->persist_grid_f32 : h->persist_grid;
  if (a.sp.halo.enabled) a.sp.halo.npush = std::min(a.sp.halo.npush, grid);
  a.ve = make_vec<T>(h, kEpiCgUpdate, kGateNone, false, 0);
  a.bar_count = h->gridbar.as<unsigned>();
  a.bar_gen = h->gridbar.as<unsigned>() + 1024;  // 4 KB away from the arrival counter: a different L2 slice
  void* params[] = {&a};
  CK(cudaLaunchCooperativeKernel((const void*)cg_persistent_kernel<T>, dim3(grid), dim3(kSpmvThreads), params,
                                 static_cast<size_t>(sizeof(T) == 4 ? h->spmv_smem_f32 : h->spmv_smem), h->stream));
  h->last_launches++;
  return 0;
}

// BiCGSTAB vector roles: x = slot X (ext), y = slot P (ext), z = slot Z (ext), p = h->yout, v = h->q, t = h->t
template <typename T>
VecArgsT<T> make_bicg_vec(b200s_handle* h, int epilogue, int gate, bool set_cond, cudaGraphConditionalHandle cond) {
  VecArgsT<T> a = make_vec<T>(h, epilogue, gate, set_cond, cond);
  a.y = slot_ptr<T>(h, kSlotP);
  a.p = h->yout.as<T>();
  return a;
}

template <typename T>
int enqueue_bicg_init(b200s_handle* h, bool set_cond, cudaGraphConditionalHandle cond) {
  int rc;
  if ((rc = launch_spmv<T>(h, slot_ptr<T>(h, kSlotX), h->t.as<T>(), nullptr, 0, kEpiNone, kGateGuess, false, 0, kSlotX)))
    return rc;
  VecArgsT<T> a = make_bicg_vec<T>(h, kEpiBiInit, kGateNone, set_cond, cond);
  LAUNCH_VEC(bicg_init_kernel<T>, a);
  return 0;
}

template <typename T>
int enqueue_bicg_body(b200s_handle* h, bool set_cond, cudaGraphConditionalHandle cond) {
  int rc;
  // re-orthogonalisation branch (BiCGSTAB.h:72-81), live only when the control state asks for it
  if ((rc = launch_spmv<T>(h, slot_ptr<T>(h, kSlotX), h->t.as<T>(), nullptr, 0, kEpiNone, kGateRestart, false, 0,
                           kSlotX)))
    return rc;
  VecArgsT<T> rs = make_bicg_vec<T>(h, kEpiBiRestart, kGateRestart, false, 0);
  LAUNCH_VEC(bicg_restart_kernel<T>, rs);
  VecArgsT<T> pa = make_bicg_vec<T>(h, kEpiNone, kGateLoop, false, 0);
  LAUNCH_VEC(bicg_p_kernel<T>, pa);
  if ((rc = launch_spmv<T>(h, slot_ptr<T>(h, kSlotP), h->q.as<T>(), h->r0.as<T>(), 1, kEpiBiR0V, kGateLoop, false, 0,
                           kSlotP)))
    return rc;
  VecArgsT<T> sa = make_bicg_vec<T>(h, kEpiNone, kGateLoop, false, 0);
  LAUNCH_VEC(bicg_s_kernel<T>, sa);
  if ((rc = launch_spmv<T>(h, slot_ptr<T>(h, kSlotZ), h->t.as<T>(), h->s.as<T>(), 2, kEpiBiTsTt, kGateLoop, false, 0,
                           kSlotZ)))
    return rc;
  VecArgsT<T> ua = make_bicg_vec<T>(h, kEpiBiUpdate, kGateLoop, set_cond, cond);
  LAUNCH_VEC(bicg_update_kernel<T>, ua);
  return 0;
}

template <typename T>
int enqueue_finalize(b200s_handle* h) {
  VecArgsT<T> a = make_vec<T>(h, kEpiNone, kGateNone, false, 0);
  LAUNCH_VEC(finalize_kernel<T>, a);
  return 0;
}

typedef int (*enqueue_fn)(b200s_handle*, bool, cudaGraphConditionalHandle);

// Builds the solve graph(s) for one solver kind.
int build_graphs(b200s_handle* h, GraphSet& g, enqueue_fn init, enqueue_fn body, int (*finalize)(b200s_handle*)) {
  if (g.built) return 0;
  int64_t saved = h->last_launches;
  if (h->loop_mode == B200S_LOOP_WHILE_GRAPH || h->loop_mode == B200S_LOOP_PERSISTENT) {
    CK(cudaGraphCreate(&g.graph, 0));
    cudaGraphConditionalHandle cond;
    CK(cudaGraphConditionalHandleCreate(&cond, g.graph, 0, cudaGraphCondAssignDefault));
    // 1. init phase, captured straight into the top-level graph
    CK(cudaStreamBeginCaptureToGraph(h->stream, g.graph, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    h->last_launches = 0;
    int rc = init(h, true, cond);
    g.init_kernels = static_cast<int>(h->last_launches);
    cudaStreamCaptureStatus st;
    const cudaGraphNode_t* deps = nullptr;
    size_t ndeps = 0;
    cudaGraph_t cap_graph = nullptr;
    if (!rc) CK(cudaStreamGetCaptureInfo_v2(h->stream, &st, nullptr, &cap_graph, &deps, &ndeps));
    std::vector<cudaGraphNode_t> dep_vec(deps, deps + ndeps);
    cudaGraph_t tmp = nullptr;
    cudaError_t ee = cudaStreamEndCapture(h->stream, &tmp);
    if (rc) return rc;
    CK(ee);
    // 2. WHILE node whose body is one iteration
    cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
    np.conditional.handle = cond;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t while_node;
    CK(cudaGraphAddNode(&while_node, g.graph, dep_vec.data(), dep_vec.size(), &np));
    cudaGraph_t body_graph = np.conditional.phGraph_out[0];
    CK(cudaStreamBeginCaptureToGraph(h->stream, body_graph, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    h->last_launches = 0;
    // several gated iterations per pass of the WHILE node: only the last one's epilogue has to set the condition
    // (a stop raised earlier leaves it untouched and simply gates the remaining kernels of the pass off)
    for (int u = 0; u < h->body_unroll && !rc; ++u) rc = body(h, true, cond);
    g.body_kernels = static_cast<int>(h->last_launches) / h->body_unroll;
    g.body_unroll = h->body_unroll;
    ee = cudaStreamEndCapture(h->stream, &tmp);
    if (rc) return rc;
    CK(ee);
    // 3. tail
    CK(cudaStreamBeginCaptureToGraph(h->stream, g.graph, &while_node, nullptr, 1, cudaStreamCaptureModeRelaxed));
    h->last_launches = 0;
    rc = finalize(h);
    g.tail_kernels = static_cast<int>(h->last_launches);
    ee = cudaStreamEndCapture(h->stream, &tmp);
    if (rc) return rc;
    CK(ee);
    CK(cudaGraphInstantiate(&g.exec, g.graph, 0));
  } else if (h->loop_mode == B200S_LOOP_CHUNKED_GRAPH) {
    const int chunk = h->cfg.chunk_iters > 0 ? h->cfg.chunk_iters : 32;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
    h->last_launches = 0;
    int rc = init(h, false, 0);
    g.init_kernels = static_cast<int>(h->last_launches);
    cudaError_t ee = cudaStreamEndCapture(h->stream, &g.graph);
    if (rc) return rc;
    CK(ee);
    CK(cudaGraphInstantiate(&g.exec, g.graph, 0));
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
    h->last_launches = 0;
    for (int i = 0; i < chunk && !rc; ++i) rc = body(h, false, 0);
    g.body_kernels = static_cast<int>(h->last_launches);
    ee = cudaStreamEndCapture(h->stream, &g.body_graph);
    if (rc) return rc;
    CK(ee);
    CK(cudaGraphInstantiate(&g.body_exec, g.body_graph, 0));
    g.tail_kernels = 1;
  } else {
    g.tail_kernels = 1;
  }
  h->last_launches = saved;
  g.built = true;
  return 0;
}

// ---- multi-column CG ------------------------------------------------------------------------------------------
// The K-wide kernels cover row-lane tiles only; matrices with two-phase or long-row tiles (strongly irregular rows),
// float factorizations and row-partitioned handles take the sequential per-column path, like the reference does.
bool multi_supported(const b200s_handle* h) {
  return h->factorized && h->precond != B200S_PRECOND_FACTORS && h->scalar_bytes == 8 && h->plan.world == 1 && h->plan.n_stream == 0 && h->plan.n_long == 0 &&
         h->spmv_impl != B200S_SPMV_DIRECT && h->plan.rows == h->plan.cols && h->plan.rows > 0;
}

template <int K>
MultiArgs<K> make_multi(b200s_handle* h, int epilogue, int gate, bool set_cond, cudaGraphConditionalHandle cond) {
  MultiArgs<K> a{};
  a.n = h->plan.rows;
  a.x = h->mx.as<double>();
  a.r = h->mr.as<double>();
  a.p = h->mp.as<double>();
  a.q = h->mq.as<double>();
  a.b = h->mb.as<double>();
  a.invdiag = h->invdiag.as<double>();
  a.S = h->mS.as<Scalars>();
  a.partials = h->mpartials.as<double>();
  a.counter = h->counter.as<unsigned>();
  a.cond_handle = static_cast<unsigned long long>(cond);
  a.set_cond = set_cond ? 1 : 0;
  a.epilogue = epilogue;
  a.gate = gate;
  return a;
}

template <int K>
int launch_spmm(b200s_handle* h, const double* x, double* y, int ndot, int epilogue, int gate) {
  SpmmArgs<K> a{};
  int rc = make_spmv_args<double>(h, a.sp, nullptr, nullptr, nullptr, kEpiNone, kGateNone, false, 0, -1);
  if (rc) return rc;
  a.m = make_multi<K>(h, epilogue, gate, false, 0);
  a.x = x;
  a.y = y;
  a.ndot = ndot;
  CK(cudaFuncSetAttribute((const void*)spmm_staged_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->spmv_smem));
  // the SAME grid as the single-column product: the grouping of the per-CTA partial sums is part of the result
  launch_k(h, spmm_staged_kernel<K>, h->spmv_grid, kSpmvThreads, static_cast<size_t>(h->spmv_smem), a);
  CK(cudaGetLastError());
  h->last_launches++;
  return 0;
}

template <int K>
int enqueue_cg_multi_init(b200s_handle* h, bool set_cond, cudaGraphConditionalHandle cond) {
  int rc;
  if ((rc = launch_spmm<K>(h, h->mx.as<double>(), h->mq.as<double>(), 0, kEpiNone, kGateGuess))) return rc;
  MultiArgs<K> a = make_multi<K>(h, kEpiCgInit, kGateNone, set_cond, cond);
  LAUNCH_VEC(cg_init_multi_kernel<K>, a);
  return 0;
}

template <int K>
int enqueue_cg_multi_body(b200s_handle* h, bool set_cond, cudaGraphConditionalHandle cond) {
  int rc;
  if ((rc = launch_spmm<K>(h, h->mp.as<double>(), h->mq.as<double>(), 1, kEpiCgPAp, kGateLoop))) return rc;
  MultiArgs<K> u = make_multi<K>(h, kEpiCgUpdate, kGateLoop, set_cond, cond);
  LAUNCH_VEC(cg_update_multi_kernel<K>, u);
  MultiArgs<K> d = make_multi<K>(h, kEpiNone, kGateNone, false, 0);
  CK(launch_k(h, cg_direction_multi_kernel<K>, h->vec_grid, kVecThreads, 0, d, h->gridbar.as<unsigned>() + 64));
  h->last_launches++;
  return 0;
}

template <int K>
int enqueue_multi_finalize(b200s_handle* h) {
  MultiArgs<K> a = make_multi<K>(h, kEpiNone, kGateNone, false, 0);
  LAUNCH_VEC(finalize_multi_kernel<K>, a);
  return 0;
}

int ensure_solver_buffers(b200s_handle* h, bool) {
  // every solver vector is a slice of the window allocation made by analyze_pattern
  if (!h->b.p || !h->r.p || !h->q.p || !h->r0.p || !h->s.p || !h->t.p || !h->yout.p)
    return fail(h, B200S_ERR_INVALID, "solve: call analyze_pattern first");
  return 0;
}

// Runs init -> loop -> finalize in the handle's loop mode.  `stop_flag` points into the pinned mirror `host` of the
// device control block(s) `dev`; the host-driven modes poll it between launches.
int drive_loop(b200s_handle* h, GraphSet& g, enqueue_fn init, enqueue_fn body, int (*finalize)(b200s_handle*),
               int (*persistent)(b200s_handle*), bool while_graph, const int* stop_flag, void* host, void* dev,
               size_t bytes) {
  int rc;
  if (persistent) {
    if ((rc = init(h, false, 0))) return rc;
    if ((rc = persistent(h))) return rc;
    return finalize(h);
  }
  if (while_graph) {
    CK(cudaGraphLaunch(g.exec, h->stream));
    return 0;
  }
  if (h->loop_mode == B200S_LOOP_CHUNKED_GRAPH) {
    CK(cudaGraphLaunch(g.exec, h->stream));
    h->last_launches += g.init_kernels;
  } else {
    if ((rc = init(h, false, 0))) return rc;
  }
  for (;;) {
    CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (*stop_flag) break;
    if (h->loop_mode == B200S_LOOP_CHUNKED_GRAPH) {
      CK(cudaGraphLaunch(g.body_exec, h->stream));
      h->last_launches += g.body_kernels;
    } else {
      if ((rc = body(h, false, 0))) return rc;
    }
  }
  return finalize(h);
}

// Row-partitioned runs: all ranks agree that everybody got this far before anything that waits on a peer is
// launched (a rank that failed on the host would otherwise leave the others spinning until the device-side timeout).
int agree_to_launch(b200s_handle* h, int my_rc) {
  if (h->plan.world <= 1 || !h->cfg.allgather) return my_rc;
  int64_t mine = my_rc;
  std::vector<int64_t> all(h->plan.world, 0);
  if (h->cfg.allgather(h->cfg.allgather_ctx, &mine, all.data(), sizeof(mine)))
    return my_rc ? my_rc : fail(h, B200S_ERR_COMM, "allgather(status before launch) failed");
  if (my_rc) return my_rc;
  for (int q = 0; q < h->plan.world; ++q)
    if (all[q] != 0)
      return fail(h, B200S_ERR_COMM, "rank " + std::to_string(q) + " failed before the launch (status " +
                                         std::to_string(all[q]) + "); nothing was launched on this rank");
  return 0;
}

int push_solve_inputs(b200s_handle* h, double tol, int64_t max_iters, int use_guess) {
  h->hS->tol = tol;
  h->hS->max_iters = max_iters;
  h->hS->use_guess = use_guess ? 1 : 0;
  h->hS->comm_error = 0;
  h->hS->comm_timeout_ns = static_cast<unsigned long long>(h->comm_timeout_ms * 1e6);
  CK(cudaMemcpyAsync(h->scalars.p, h->hS, kSolveInputBytes, cudaMemcpyHostToDevice, h->stream));
  return 0;
}

template <typename T>
int run_solve(b200s_handle* h, bool bicg, const T* b_dev, T* x_dev, int use_guess, double tol, int64_t max_iters,
              int64_t* iters_out, double* error_out, int* info_out) {
  int rc = 0;
  if (!h->factorized)
    rc = fail(h, B200S_ERR_INVALID, "solve: call analyze_pattern + factorize first (IterativeSolverBase.h:337 asserts m_isInitialized)");
  else if (h->scalar_bytes != static_cast<int>(sizeof(T)))
    rc = fail(h, B200S_ERR_INVALID, "solve: scalar type differs from the one given to factorize");
  else if (h->plan.world == 1 && h->plan.rows != h->plan.cols)
    rc = fail(h, B200S_ERR_INVALID, "solve: the matrix is not square");
  GraphSet& g = bicg ? h->bicg : h->cg;
  if (!rc) rc = ensure_solver_buffers(h, bicg);
  if (!rc)
    rc = build_graphs(h, g, bicg ? enqueue_bicg_init<T> : enqueue_cg_init<T>, bicg ? enqueue_bicg_body<T> : enqueue_cg_body<T>,
                      enqueue_finalize<T>);
  if ((rc = agree_to_launch(h, rc))) return rc;

  const int64_t n = h->plan.rows;
  if (tol < 0) tol = std::numeric_limits<T>::epsilon();       // IterativeSolverBase.h:413
  if (sizeof(T) == 4) tol = static_cast<double>(static_cast<float>(tol));  // RealScalar m_tolerance
  if (max_iters < 0) max_iters = 2 * h->plan.cols;             // :281-284
  T* x_int = slot_ptr<T>(h, kSlotX);
  if (b_dev != h->b.as<T>()) CK(cudaMemcpyAsync(h->b.p, b_dev, n * sizeof(T), cudaMemcpyDeviceToDevice, h->stream));
  if (use_guess && x_dev != x_int) CK(cudaMemcpyAsync(x_int, x_dev, n * sizeof(T), cudaMemcpyDeviceToDevice, h->stream));
  if ((rc = push_solve_inputs(h, tol, max_iters, use_guess))) return rc;

  h->last_launches = 0;
  CK(cudaEventRecord(h->ev0, h->stream));
  const bool persistent = (h->loop_mode == B200S_LOOP_PERSISTENT) && !bicg && h->spmv_impl != B200S_SPMV_DIRECT;
  const bool while_graph = (h->loop_mode == B200S_LOOP_WHILE_GRAPH) || (h->loop_mode == B200S_LOOP_PERSISTENT && !persistent);
  if ((rc = drive_loop(h, g, bicg ? enqueue_bicg_init<T> : enqueue_cg_init<T>, bicg ? enqueue_bicg_body<T> : enqueue_cg_body<T>,
                       enqueue_finalize<T>, persistent ? launch_cg_persistent<T> : nullptr, while_graph, &h->hS->stop,
                       h->hS, h->scalars.p, sizeof(Scalars))))
    return rc;
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaMemcpyAsync(h->hS, h->scalars.p, sizeof(Scalars), cudaMemcpyDeviceToHost, h->stream));
  if (x_dev != x_int) CK(cudaMemcpyAsync(x_dev, x_int, n * sizeof(T), cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->last_solve_ms = ms;

  const Scalars& S = *h->hS;
  h->last_comm_error = S.comm_error;
  h->last_nonfinite = S.numerical_issue;
  h->last_restarts = bicg ? S.restarts : 0;
  if (S.comm_error) {
    if (info_out) *info_out = 1;  // NumericalIssue: the result is not usable
    return fail(h, B200S_ERR_COMM, "a device-side wait on a peer GPU expired after " + std::to_string(h->comm_timeout_ms) +
                                       " ms (dead or diverged rank); the solve was abandoned");
  }
  int64_t iters;
  double err;
  auto real_sqrt_ratio = [](double rr, double bb) {  // tol_error = sqrt(residualNorm2 / rhsNorm2) in RealScalar
    if (sizeof(T) == 4) return static_cast<double>(std::sqrt(static_cast<float>(rr) / static_cast<float>(bb)));
    return std::sqrt(rr / bb);
  };
  if (bicg && S.rhs_zero) {  // BiCGSTAB.h:47-51 returns before touching iters / tol_error
    iters = max_iters;
    err = tol;
  } else if (S.rhs_zero) {   // ConjugateGradient.h:46-52
    iters = 0;
    err = 0.0;
  } else if (!bicg && S.numerical_issue) {
    // CG on a non-finite residual: the reference iterates on NaNs until maxIters and reports that (see kEpiCgInit)
    iters = max_iters;
    err = std::numeric_limits<double>::quiet_NaN();
  } else {
    iters = S.iter;
    err = real_sqrt_ratio(S.rr, S.bb);  // ConjugateGradient.h:89 / BiCGSTAB.h:104
  }
  int info = (err <= tol) ? 0 : 2;  // ConjugateGradient.h:220 / BiCGSTAB.h:201-203 (NaN compares false: NoConvergence)
  if (iters_out) *iters_out = iters;
  if (error_out) *error_out = err;
  if (info_out) *info_out = info;
  h->last_iterations = iters;
  h->last_spmv = S.spmv_count;
  if (while_graph) {
    // body executions: one per SpMV launched inside the loop (CG: 1 per pass; BiCGSTAB: 2 per pass + restarts)
    int64_t init_spmv = use_guess ? 1 : 0;
    if (bicg) init_spmv = 1;  // counted by kEpiBiInit regardless (the gated launch still happens)
    int64_t loop_spmv = std::max<int64_t>(0, S.spmv_count - init_spmv);
    int64_t passes = bicg ? (loop_spmv - S.restarts + 1) / 2 : loop_spmv;
    // the WHILE node runs whole passes of body_unroll iterations; kernels of iterations past the stop are launched
    // too (they return at their gate)
    int64_t wpasses = (passes + g.body_unroll - 1) / g.body_unroll;
    h->last_launches = g.init_kernels + wpasses * g.body_unroll * g.body_kernels + g.tail_kernels;
  }
  return 0;
}

// One batch of K (2, 4 or 8) columns, device-resident column-major B / X (leading dimensions ldb / ldx >= rows).
template <int K>
int run_solve_multi(b200s_handle* h, int ncols, const double* B_dev, int64_t ldb, double* X_dev, int64_t ldx,
                    int use_guess, double tol, int64_t max_iters, int64_t* iters_out, double* error_out,
                    int* info_out) {
  constexpr int LOGK = (K == 2) ? 1 : (K == 4) ? 2 : 3;
  const int64_t n = h->plan.rows;
  const size_t vb = static_cast<size_t>(n) * K * 8;
  int rc;
  if ((rc = dev_alloc(h, h->mx, vb))) return rc;
  if ((rc = dev_alloc(h, h->mr, vb))) return rc;
  if ((rc = dev_alloc(h, h->mp, vb))) return rc;
  if ((rc = dev_alloc(h, h->mq, vb))) return rc;
  if ((rc = dev_alloc(h, h->mb, vb))) return rc;
  if ((rc = dev_alloc(h, h->mS, sizeof(Scalars) * kMultiMax, false))) return rc;
  if ((rc = dev_alloc(h, h->mpartials, sizeof(double) * kMaxGrid * kMultiPartialStride, false))) return rc;
  if (!h->hSm) CK(cudaMallocHost(reinterpret_cast<void**>(&h->hSm), sizeof(Scalars) * kMultiMax));
  GraphSet& g = h->cg_multi[LOGK];
  if (g.built && (g.tag != h->mx.p)) destroy_graphs(g);  // the vectors were reallocated: graphs bake pointers
  if ((rc = build_graphs(h, g, enqueue_cg_multi_init<K>, enqueue_cg_multi_body<K>, enqueue_multi_finalize<K>))) return rc;
  g.tag = h->mx.p;

  if (tol < 0) tol = std::numeric_limits<double>::epsilon();
  if (max_iters < 0) max_iters = 2 * h->plan.cols;
  const int ig = std::max(1, std::min(h->sm_count * 8, static_cast<int>((n + 255) / 256)));
  interleave_kernel<K><<<ig, 256, 0, h->stream>>>(n, ncols, B_dev, ldb, h->mb.as<double>());
  if (use_guess) interleave_kernel<K><<<ig, 256, 0, h->stream>>>(n, ncols, X_dev, ldx, h->mx.as<double>());
  CK(cudaGetLastError());
  std::memset(h->hSm, 0, sizeof(Scalars) * kMultiMax);
  for (int j = 0; j < K; ++j) {
    h->hSm[j].tol = tol;
    h->hSm[j].max_iters = max_iters;
    h->hSm[j].use_guess = use_guess ? 1 : 0;
  }
  CK(cudaMemcpyAsync(h->mS.p, h->hSm, sizeof(Scalars) * kMultiMax, cudaMemcpyHostToDevice, h->stream));

  h->last_launches = 0;
  CK(cudaEventRecord(h->ev0, h->stream));
  const bool while_graph = (h->loop_mode == B200S_LOOP_WHILE_GRAPH) || (h->loop_mode == B200S_LOOP_PERSISTENT);
  if ((rc = drive_loop(h, g, enqueue_cg_multi_init<K>, enqueue_cg_multi_body<K>, enqueue_multi_finalize<K>, nullptr,
                       while_graph, &h->hSm[0].stop_all, h->hSm, h->mS.p, sizeof(Scalars) * kMultiMax)))
    return rc;
  CK(cudaEventRecord(h->ev1, h->stream));
  deinterleave_kernel<K><<<ig, 256, 0, h->stream>>>(n, ncols, h->mx.as<double>(), X_dev, ldx);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h->hSm, h->mS.p, sizeof(Scalars) * kMultiMax, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->last_solve_ms += ms;
  int64_t max_spmv = 0;
  for (int j = 0; j < ncols; ++j) {
    const Scalars& S = h->hSm[j];
    int64_t iters;
    double err;
    if (S.rhs_zero) { iters = 0; err = 0.0; }                       // ConjugateGradient.h:46-52
    else if (S.numerical_issue) { iters = max_iters; err = std::numeric_limits<double>::quiet_NaN(); }
    else { iters = S.iter; err = std::sqrt(S.rr / S.bb); }          // :89
    if (iters_out) iters_out[j] = iters;
    if (error_out) error_out[j] = err;
    if (info_out) info_out[j] = (err <= tol) ? 0 : 2;               // :220
    max_spmv = std::max<int64_t>(max_spmv, S.spmv_count);
    h->last_iterations = iters;
    if (S.numerical_issue) h->last_nonfinite = S.numerical_issue;
  }
  h->last_spmv += max_spmv;
  if (while_graph) {
    const int64_t passes = std::max<int64_t>(0, max_spmv - (use_guess ? 1 : 0));
    const int64_t wpasses = (passes + g.body_unroll - 1) / g.body_unroll;
    h->last_launches = g.init_kernels + wpasses * g.body_unroll * g.body_kernels + g.tail_kernels;
  }
  return 0;
}

// Any number of columns: batches of 8 / 4 / 2 through the K-wide kernels, a last single column (or everything, when
// the handle does not support batching) through the single-column solver.  Column-major, device pointers.
int cg_general_solve(b200s_handle* h, const double* b_src, double* x_dst, bool device_ptrs, int use_guess, double tol,
                     int64_t max_iters, int64_t* iters_out, double* error_out, int* info_out);  // precond.inc

int solve_multi_device(b200s_handle* h, int64_t ncols, const double* B, int64_t ldb, double* X, int64_t ldx,
                       int use_guess, double tol, int64_t max_iters, int64_t* iters_out, double* error_out,
                       int* info_out, int64_t* launches_out) {
  const bool batched = multi_supported(h) && env_int("B200S_MULTI_RHS", 1) != 0;
  double total_ms = 0;
  int64_t launches = 0;
  h->last_spmv = 0;
  h->last_nonfinite = 0;
  int64_t c = 0;
  while (c < ncols) {
    const int64_t left = ncols - c;
    int64_t *it = iters_out ? iters_out + c : nullptr;
    double* er = error_out ? error_out + c : nullptr;
    int* in = info_out ? info_out + c : nullptr;
    int rc, took;
    h->last_solve_ms = 0;
    if (batched && left >= 2) {
      // Measured at 256^3 (profiles/r2_multi_rhs.jsonl): 4 columns per batch run at 0.97 of the HBM roofline of the
      // batched iteration (1.79x the per-column rate of single solves); 8 per batch are 20 % slower per column
      // (register pressure, L1 wavefronts of 64-byte rows), so wider blocks are cut into batches of 4.
      const int kmax = std::max(2, std::min(kMultiMax, env_int("B200S_MULTI_K", 4)));
      if (left >= 5 && kmax >= 8) { took = static_cast<int>(std::min<int64_t>(8, left)); rc = run_solve_multi<8>(h, took, B + c * ldb, ldb, X + c * ldx, ldx, use_guess, tol, max_iters, it, er, in); }
      else if (left >= 3 && kmax >= 4) { took = static_cast<int>(std::min<int64_t>(4, left)); rc = run_solve_multi<4>(h, took, B + c * ldb, ldb, X + c * ldx, ldx, use_guess, tol, max_iters, it, er, in); }
      else { took = 2; rc = run_solve_multi<2>(h, took, B + c * ldb, ldb, X + c * ldx, ldx, use_guess, tol, max_iters, it, er, in); }
    } else {
      took = 1;
      const int64_t spmv_before = h->last_spmv;
      if (h->precond == B200S_PRECOND_FACTORS)
        rc = cg_general_solve(h, B + c * ldb, X + c * ldx, true, use_guess, tol, max_iters, it, er, in);
      else
        rc = run_solve<double>(h, false, B + c * ldb, X + c * ldx, use_guess, tol, max_iters, it, er, in);
      h->last_spmv += spmv_before;
    }
    if (rc) return rc;
    total_ms += h->last_solve_ms;
    launches += h->last_launches;
    c += took;
  }
  h->last_solve_ms = total_ms;
  h->last_launches = launches;
  if (launches_out) *launches_out = launches;
  return 0;
}

template <typename T>
int factorize_impl(b200s_handle* h, const T* values, int precond) {
  if (!h->analyzed) return fail(h, B200S_ERR_INVALID, "factorize: call analyze_pattern first (IterativeSolverBase.h:218 asserts m_analysisIsOk)");
  const Plan& p = h->plan;
  int rc;
  // Row-partitioned one-triangle input: the mirror images of entries stored on other ranks arrive through the
  // config's allgather (setup path, host memory).  A collective: entered before any early return that depends on
  // this rank's arguments alone, so that a bad argument on one rank cannot leave the others waiting.
  std::vector<unsigned char> imports;
  if (!p.tri_counts.empty()) {
    std::vector<unsigned char> zeros;
    const void* src_vals = values;
    if (!values) { zeros.assign(static_cast<size_t>(std::max<int64_t>(p.input_nnz, 1)) * sizeof(T), 0); src_vals = zeros.data(); }
    std::string err;
    if ((rc = exchange_mirror_values(h->cfg, p, src_vals, sizeof(T), imports, err))) return fail(h, rc, err);
  }
  if (!values && p.input_nnz > 0) return fail(h, B200S_ERR_INVALID, "factorize: values is null");
  if (precond != B200S_PRECOND_IDENTITY && precond != B200S_PRECOND_JACOBI)
    return fail(h, B200S_ERR_INVALID, "factorize: precond must be 0 (identity) or 1 (Jacobi)");
  CK(cudaSetDevice(h->device));
  if ((rc = dev_alloc(h, h->vals, static_cast<size_t>(p.nnz) * sizeof(T) + 64))) return rc;
  if (h->scalar_bytes != static_cast<int>(sizeof(T))) {  // graphs bake pointers/types: rebuild lazily
    destroy_graphs(h->cg);
    destroy_graphs(h->bicg);
    for (GraphSet& g : h->cg_multi) destroy_graphs(g);
  }
  h->scalar_bytes = sizeof(T);
  h->precond = h->precond_factorize = precond;
  free_factors(h);  // factors of the previous values are stale (IterativeSolverBase::factorize re-factorizes, :216-224)
  cudaEvent_t e0 = h->ev0, e1 = h->ev1;
  CK(cudaEventRecord(e0, h->stream));
  if (p.src.empty()) {
    if (p.nnz) CK(cudaMemcpyAsync(h->vals.p, values, static_cast<size_t>(p.nnz) * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  } else {
    DevBuf staging;
    if ((rc = dev_alloc(h, staging, static_cast<size_t>(p.input_nnz + p.n_import) * sizeof(T) + 64))) return rc;
    if (p.input_nnz) CK(cudaMemcpyAsync(staging.p, values, static_cast<size_t>(p.input_nnz) * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    if (!imports.empty())  // mirrors stored on other ranks sit behind the caller's own values (plan.h: src >= input_nnz)
      CK(cudaMemcpyAsync(staging.as<T>() + p.input_nnz, imports.data(), imports.size(), cudaMemcpyHostToDevice, h->stream));
    if (p.nnz) {
      gather_values_kernel<T><<<h->sm_count * 4, 256, 0, h->stream>>>(p.nnz, h->src.as<int32_t>(), staging.as<T>(), h->vals.as<T>());
      CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(h->stream));
    dev_free(h, staging);
  }
  if (p.rows) {
    // the diagonal of local row j sits at local column j (owned columns are numbered like the rows)
    jacobi_factorize_kernel<T><<<static_cast<int>((p.rows + 255) / 256), 256, 0, h->stream>>>(
        static_cast<int>(p.rows), h->rowptr.as<int32_t>(), h->colidx.as<int32_t>(), h->vals.as<T>(), h->invdiag.as<T>(),
        precond == B200S_PRECOND_IDENTITY);
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(e1, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  h->last_h2d_ms = ms;
  h->factorized = true;
  return 0;
}

template <typename T>
int spmv_device_impl(b200s_handle* h, const T* x_dev, T* y_dev, int reps, float* ms_avg) {
  int rc = 0;
  if (!h->factorized) rc = fail(h, B200S_ERR_INVALID, "spmv: call analyze_pattern + factorize first");
  else if (h->scalar_bytes != static_cast<int>(sizeof(T))) rc = fail(h, B200S_ERR_INVALID, "spmv: scalar type differs from factorize");
  else if (!x_dev || !y_dev) rc = fail(h, B200S_ERR_INVALID, "spmv: null pointer");
  if ((rc = agree_to_launch(h, rc))) return rc;
  CK(cudaSetDevice(h->device));
  if (reps < 1) reps = 1;
  const T* x_ext = x_dev;
  h->last_launches = 0;
  if (h->plan.world > 1) {  // the product reads x from the peer-visible window: [owned | ghost]
    T* xs = slot_ptr<T>(h, kSlotSpmv);
    if (x_dev != xs) CK(cudaMemcpyAsync(xs, x_dev, static_cast<size_t>(h->plan.rows) * sizeof(T), cudaMemcpyDeviceToDevice, h->stream));
    x_ext = xs;
    if ((rc = push_solve_inputs(h, 0.0, 0, 0))) return rc;  // clears comm_error, sets the wait bound
  }
  CK(cudaEventRecord(h->ev0, h->stream));
  for (int i = 0; i < reps; ++i) {
    // multi-GPU: each launch pushes the halo, multiplies, and ends in a cross-rank rendezvous (kEpiSpmvOnly) that
    // keeps the single-buffered ghost slots safe for the next push
    if ((rc = launch_spmv<T>(h, x_ext, y_dev, nullptr, 0, kEpiNone, kGateNone, false, 0, kSlotSpmv))) return rc;
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  if (h->plan.world > 1) CK(cudaMemcpyAsync(h->hS, h->scalars.p, kSolveInputBytes, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  if (ms_avg) *ms_avg = ms / reps;
  h->last_solve_ms = ms;
  if (h->plan.world > 1 && h->hS->comm_error) {
    h->last_comm_error = 1;
    return fail(h, B200S_ERR_COMM, "spmv: a device-side wait on a peer GPU expired (dead or diverged rank)");
  }
  return 0;
}

template <typename T>
int spmv_host_impl(b200s_handle* h, const T* x, T* y) {
  if (!h->factorized) return fail(h, B200S_ERR_INVALID, "spmv: call analyze_pattern + factorize first");
  if (!x || !y) return fail(h, B200S_ERR_INVALID, "spmv: null pointer");
  CK(cudaSetDevice(h->device));
  const Plan& p = h->plan;
  const int64_t nx = (p.world == 1) ? p.cols : p.rows;
  T* xs = slot_ptr<T>(h, kSlotSpmv);
  int rc;
  CK(cudaEventRecord(h->ev0, h->stream));
  CK(cudaMemcpyAsync(xs, x, static_cast<size_t>(nx) * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->last_h2d_ms = ms;
  float k_ms = 0;
  if ((rc = spmv_device_impl<T>(h, xs, h->yout.as<T>(), 1, &k_ms))) return rc;
  CK(cudaEventRecord(h->ev0, h->stream));
  CK(cudaMemcpyAsync(y, h->yout.p, static_cast<size_t>(p.rows) * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->last_d2h_ms = ms;
  h->last_solve_ms = k_ms;
  return 0;
}

template <typename T>
int solve_host(b200s_handle* h, bool bicg, const T* b, T* x, int use_guess, double tol, int64_t max_iters,
               int64_t* iters_out, double* error_out, int* info_out) {
  if (!h) return B200S_ERR_INVALID;
  if (!b || !x) return fail(h, B200S_ERR_INVALID, "solve: null pointer");
  CK(cudaSetDevice(h->device));
  int rc = 0;
  if (!h->factorized) rc = fail(h, B200S_ERR_INVALID, "solve: call analyze_pattern + factorize first");
  if (!rc) rc = ensure_solver_buffers(h, bicg);
  if (rc) return agree_to_launch(h, rc);  // let the peers know instead of leaving them waiting
  const size_t vb = static_cast<size_t>(h->plan.rows) * sizeof(T);
  cudaEvent_t e0 = h->ev0, e1 = h->ev1;
  CK(cudaEventRecord(e0, h->stream));
  CK(cudaMemcpyAsync(h->b.p, b, vb, cudaMemcpyHostToDevice, h->stream));
  if (use_guess) CK(cudaMemcpyAsync(slot_ptr<T>(h, kSlotX), x, vb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaEventRecord(e1, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  h->last_h2d_ms = ms;
  if ((rc = run_solve<T>(h, bicg, h->b.as<T>(), slot_ptr<T>(h, kSlotX), use_guess, tol, max_iters, iters_out, error_out,
                         info_out)))
    return rc;
  double solve_ms = h->last_solve_ms;
  CK(cudaEventRecord(e0, h->stream));
  CK(cudaMemcpyAsync(x, slot_ptr<T>(h, kSlotX), vb, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaEventRecord(e1, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  h->last_d2h_ms = ms;
  h->last_solve_ms = solve_ms;
  return 0;
}

template <typename T>
int solve_device(b200s_handle* h, bool bicg, const T* b, T* x, int use_guess, double tol, int64_t max_iters,
                 int64_t* iters_out, double* error_out, int* info_out) {
  if (!h) return B200S_ERR_INVALID;
  if (!b || !x) return fail(h, B200S_ERR_INVALID, "solve: null pointer");
  CK(cudaSetDevice(h->device));
  return run_solve<T>(h, bicg, b, x, use_guess, tol, max_iters, iters_out, error_out, info_out);
}

int configure_l2_persistence(b200s_handle* h);

int configure_spmv(b200s_handle* h) {
  // stage geometry is decided for the widest scalar (double) so that one plan serves both precisions
  const Plan& p = h->plan;
  size_t stage = spmv_stage_bytes<double>(p.tile_nnz, p.tile_rows_cap);
  // Two stages per CTA and as many CTAs per SM as shared memory allows (four with the default 2048-nnz tile):
  // measured on B200 at 256^3, 4 CTAs x 2 stages reaches 93 % of the copy bandwidth where 2 CTAs x 4 stages
  // reaches 65 % -- the x gathers need the extra resident warps (profiles/r1_spmv_geometry_sweep.jsonl).
  int stages = env_int("B200S_SPMV_STAGES", 2);
  stages = std::max(2, std::min(8, stages));
  size_t budget = static_cast<size_t>(env_int("B200S_SPMV_SMEM_KB", 56)) * 1024;  // per CTA
  while (stages > 2 && stage * stages > budget) --stages;
  if (stage * stages > 220 * 1024) return fail(h, B200S_ERR_INVALID, "tile_nnz/tile_rows too large for shared memory");
  h->spmv_stages = stages;
  h->spmv_smem = static_cast<int>(stage * stages);
  // float tiles hold two thirds of the bytes of double tiles.  A third stage (same bytes in flight per CTA as double)
  // was measured and rejected: it costs two of the six resident CTAs per SM and the 7-point product drops from 0.816 to
  // 0.750 of the HBM figure (profiles/r2_exp_l2_persist_f32_stages.txt) -- resident warps matter more than depth.
  const size_t stage32 = spmv_stage_bytes<float>(p.tile_nnz, p.tile_rows_cap);
  int stages32 = std::max(2, std::min(8, env_int("B200S_SPMV_STAGES_F32", stages)));
  while (stages32 > 2 && stage32 * stages32 > std::max(budget, stage * stages)) --stages32;
  h->spmv_stages_f32 = stages32;
  h->spmv_smem_f32 = static_cast<int>(stage32 * stages32);
  const void* fns[] = {(const void*)spmv_staged_kernel<double, 0>, (const void*)spmv_staged_kernel<double, 1>,
                       (const void*)spmv_staged_kernel<double, 2>};
  const void* fns32[] = {(const void*)spmv_staged_kernel<float, 0>, (const void*)spmv_staged_kernel<float, 1>,
                         (const void*)spmv_staged_kernel<float, 2>};
  for (const void* f : fns) CK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, h->spmv_smem));
  for (const void* f : fns32) CK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, h->spmv_smem_f32));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmv_staged_kernel<double, 1>, kSpmvThreads, h->spmv_smem));
  if (occ < 1) return fail(h, B200S_ERR_CUDA, "staged SpMV kernel does not fit on an SM");
  occ = std::min(occ, env_int("B200S_SPMV_OCC", 8));
  int grid = h->sm_count * occ;
  int ntiles = static_cast<int>(p.tiles.size());
  grid = std::max(1, std::min(grid, std::max(1, ntiles)));
  h->spmv_grid = std::min(grid, kMaxGrid);
  {
    int occ32 = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ32, spmv_staged_kernel<float, 1>, kSpmvThreads, h->spmv_smem_f32));
    occ32 = std::max(1, std::min(occ32, env_int("B200S_SPMV_OCC", 8)));
    h->spmv_grid_f32 = std::min(std::max(1, std::min(h->sm_count * occ32, std::max(1, ntiles))), kMaxGrid);
  }
  {
    CK(cudaFuncSetAttribute((const void*)cg_persistent_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->spmv_smem));
    CK(cudaFuncSetAttribute((const void*)cg_persistent_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->spmv_smem_f32));
    int pocc = 0, pocc32 = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pocc, cg_persistent_kernel<double>, kSpmvThreads, h->spmv_smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pocc32, cg_persistent_kernel<float>, kSpmvThreads, h->spmv_smem_f32));
    pocc = std::min(pocc, env_int("B200S_SPMV_OCC", 8));
    pocc32 = std::min(pocc32, env_int("B200S_SPMV_OCC", 8));
    h->persist_grid = std::max(1, std::min(h->sm_count * std::max(1, pocc), kMaxGrid));
    h->persist_grid_f32 = std::max(1, std::min(h->sm_count * std::max(1, pocc32), kMaxGrid));
  }
  int64_t n2 = std::max<int64_t>(1, p.rows / 2);
  // CTAs per SM of the fused vector kernels: 6 for HBM-sized vectors; 4 below ~4M rows, where the vectors are (partly)
  // L2-resident and the ticket / fold tail of the reduction weighs more than bytes in flight (profiles/r2_exp_vec_grid.txt:
  // 128^3 73.1 -> 70.5 us per iteration, 2D 1024^2 46.3 -> 42.5; 256^3 503 -> 535 the other way).
  const int vec_ctas = env_int("B200S_VEC_CTAS_PER_SM", p.rows < 4000000 ? 4 : 6);
  int64_t vg = std::min<int64_t>(static_cast<int64_t>(h->sm_count) * vec_ctas, (n2 + kVecThreads - 1) / kVecThreads);
  h->vec_grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(vg, kMaxGrid)));
  // L2 policy of the matrix stream (TMA cache hint).  The matrix is read once per product; everything else (x gathers,
  // the vectors the next kernels read) profits from staying in L2.  Measured:
  //   * one GPU, vectors fit L2 but the matrix does not (128^3): evict-first 83.5 -> 73.1 us per CG iteration;
  //   * one GPU, vectors far larger than L2 (256^3): neutral (-0.6 %), so the default policy is kept there;
  //   * row-partitioned runs (peer-mapped windows), 16 M rows per rank: 606 -> 544 us per iteration with evict-first,
  //     SpMV alone 0.305 -> 0.274 ms -- always on for world > 1;
  //   * everything fits L2 (2D 1024^2): default policy, the matrix should stay resident too.
  if (env_int("B200S_EVICT_FIRST", -1) < 0) {
    int l2 = 0;
    CK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, h->device));
    const double vec_bytes = 6.0 * 8.0 * static_cast<double>(p.rows + static_cast<int64_t>(p.ghost_cols.size()));
    const double mat_bytes = 12.0 * static_cast<double>(p.nnz) + 4.0 * static_cast<double>(p.rows);
    const bool all_fits = l2 > 0 && vec_bytes + mat_bytes <= 1.0 * l2;
    if (p.world > 1)
      h->evict_first = all_fits ? 0 : 1;
    else
      h->evict_first = (l2 > 0 && vec_bytes <= 1.75 * l2 && !all_fits) ? 1 : 0;
  }
  // Early x update (cg_direction_kernel): it splits the direction pass in two, so p is read twice.  That is free while
  // the CG vectors live in L2 (strong scaling: 256^3 over 8 GPUs = 2.1 M rows per GPU; one GPU 128^3: 76.7 -> 74.8 us per
  // iteration) and costs a full extra HBM pass otherwise (one GPU 256^3: 503 -> 539 us), so it is tied to the size.
  {
    int l2 = 0;
    CK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, h->device));
    const double six_vectors = 6.0 * 8.0 * static_cast<double>(p.rows + static_cast<int64_t>(p.ghost_cols.size()));
    const int want = env_int("B200S_EARLY_X", -1);
    h->early_x = want >= 0 ? want : (l2 > 0 && six_vectors <= 1.0 * l2 ? 1 : 0);
  }
  // AUTO: the WHILE graph everywhere.  Round 1 picked
  // the persistent kernel for >= 4M rows on one GPU as well (524 vs 544 us per iteration at 256^3); since the SpMV
  // fetches its tile descriptors a round ahead the stand-alone kernel runs at 250 us (1.06 of the measured copy
  // bandwidth) and the graph wins: 500.6 vs 516.3 us per iteration, same box (profiles/r2_exp_descriptor_prefetch.txt).
  if (h->loop_auto) {
    // ... and on 8 GPUs the graph wins as well since then: 256^3 11,740 vs 10,890 it/s, 512^3 1,963 vs 1,879 it/s
    // (profiles/r2_exp_loop_mode_x8.txt).  The persistent kernel stays available (B200S_LOOP_PERSISTENT).
    const int mg = env_int("B200S_LOOP_MODE_MULTI", B200S_LOOP_WHILE_GRAPH);
    h->loop_mode = (p.world > 1) ? mg : B200S_LOOP_WHILE_GRAPH;
  }
  // direct kernel: lanes per row from the global mean row length
  int mean = p.rows ? static_cast<int>((p.nnz + p.rows - 1) / p.rows) : 1;
  int lg = 0;
  while ((1 << lg) < std::max(1, mean / 2) && lg < 5) ++lg;
  h->direct_lg = lg;
  return 0;
}

// L2 residency of the CG working set.  When the five vectors of a CG iteration (r, Ap, D^-1, x, p: the front of the
// window) fit the persisting share of L2 while the matrix does not fit L2, they are marked persisting with a stream
// access-policy window (captured into the graph's kernel nodes as well); the matrix stream already carries an
// evict-first hint, so one iteration re-reads only the matrix from HBM.  This is the regime of strong scaling over
// 8 GPUs (256^3 / 8 = 2.1 M rows per GPU: 84 MB of vectors, 176 MB of matrix, 126 MB of L2).
int configure_l2_persistence(b200s_handle* h) {
  h->l2_persist = 0;
  cudaStreamAttrValue attr;
  std::memset(&attr, 0, sizeof(attr));
  const int want = env_int("B200S_L2_PERSIST", -1);  // -1 auto, 0 off, 1 force
  int l2 = 0, max_persist = 0, max_window = 0;
  CK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, h->device));
  CK(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, h->device));
  CK(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, h->device));
  const Plan& p = h->plan;
  const size_t hot = h->slots_off + 2 * h->slot_bytes;  // r, q, invdiag, slot X, slot P
  const double mat_bytes = 12.0 * static_cast<double>(p.nnz) + 4.0 * static_cast<double>(p.rows);
  bool on = want > 0;
  if (want < 0)
    on = max_persist > 0 && hot <= static_cast<size_t>(max_persist) && hot <= static_cast<size_t>(max_window) &&
         mat_bytes + static_cast<double>(hot) > 1.0 * l2;
  if (on && max_persist > 0 && max_window > 0) {
    const size_t bytes = std::min(hot, static_cast<size_t>(max_window));
    const size_t setaside = std::min(static_cast<size_t>(max_persist),
                                     static_cast<size_t>(env_int("B200S_L2_PERSIST_MB", 0)) > 0
                                         ? static_cast<size_t>(env_int("B200S_L2_PERSIST_MB", 0)) << 20
                                         : bytes);
    CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setaside));
    attr.accessPolicyWindow.base_ptr = h->window.p;
    attr.accessPolicyWindow.num_bytes = bytes;
    attr.accessPolicyWindow.hitRatio = static_cast<float>(std::min(1.0, static_cast<double>(setaside) / static_cast<double>(bytes)));
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    h->l2_persist = 1;
  }
  CK(cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
  return 0;
}

int k_precond(b200s_handle* h, int64_t n, const double* r, double* z);  // precond.inc
#include "krylov.inc"
#include "precond.inc"

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int b200s_version(void) { return B200S_VERSION; }

int b200s_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
  }
  return ok;
}

const char* b200s_last_error(const b200s_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int b200s_create(const b200s_config* cfg, b200s_handle** out) {
  if (!out) return B200S_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    g_create_error = "no CUDA device: this library has no CPU path (sm_100a kernels only)";
    return B200S_ERR_NO_DEVICE;
  }
  b200s_handle* h = new b200s_handle();
  h->cfg.struct_size = sizeof(b200s_config);
  h->cfg.device = -1;
  h->cfg.world = 1;
  if (cfg) std::memcpy(&h->cfg, cfg, std::min<size_t>(sizeof(b200s_config), cfg->struct_size > 0 ? cfg->struct_size : sizeof(b200s_config)));
  if (h->cfg.world <= 0) h->cfg.world = 1;
  auto bail = [&](int code, const std::string& m) {
    g_create_error = m;
    delete h;
    return code;
  };
  if (h->cfg.world > kMaxWorld) return bail(B200S_ERR_UNSUPPORTED, "world > 8 (one NVSwitch box)");
  int dev = h->cfg.device;
  if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) return bail(B200S_ERR_CUDA, "cudaGetDevice failed");
  if (dev >= ndev) return bail(B200S_ERR_INVALID, "device ordinal out of range");
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10)
    return bail(B200S_ERR_NO_DEVICE, "device " + std::to_string(dev) + " is sm_" + std::to_string(major) + std::to_string(minor) +
                                         "; this library ships sm_100a code only and has no fallback");
  h->device = dev;
  if (cudaSetDevice(dev) != cudaSuccess) return bail(B200S_ERR_CUDA, "cudaSetDevice failed");
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess ||
      cudaMallocHost(reinterpret_cast<void**>(&h->hS), sizeof(Scalars)) != cudaSuccess) {
    std::string m = std::string("stream/event creation failed: ") + cudaGetErrorString(cudaGetLastError());
    b200s_destroy(h);
    g_create_error = m;
    return B200S_ERR_CUDA;
  }
  std::memset(h->hS, 0, sizeof(Scalars));
  // AUTO resolves to the WHILE graph (see configure_spmv for the measurements behind that; round 1 resolved to the
  // persistent cooperative kernel for row-partitioned and large problems, profiles/r1_loop_overheads.txt).
  h->loop_mode = h->cfg.loop_mode ? h->cfg.loop_mode : env_int("B200S_LOOP_MODE", B200S_LOOP_AUTO);
  h->loop_auto = (h->loop_mode == B200S_LOOP_AUTO);
  if (h->loop_auto) h->loop_mode = B200S_LOOP_WHILE_GRAPH;  // refined per problem in configure_spmv
  h->spmv_impl = h->cfg.spmv_impl ? h->cfg.spmv_impl : env_int("B200S_SPMV_IMPL", B200S_SPMV_STAGED);
  h->evict_first = env_int("B200S_EVICT_FIRST", -1);  // -1: decided per problem in configure_spmv
  // Programmatic dependent launch between the solver kernels.  Round 1's early trigger (at kernel entry) cost 7 % at
  // 256^3 and was off; the trigger now fires when a CTA's streaming work is done, which is neutral by itself
  // (profiles/r2_exp_pdl.txt) and lets cg_direction apply x += alpha p behind cg_update's cross-rank all-reduce.
  h->pdl = env_int("B200S_PDL", 1);
  h->body_unroll = std::max(1, std::min(16, env_int("B200S_BODY_UNROLL", 4)));
  h->comm_timeout_ms = std::max(1, env_int("B200S_COMM_TIMEOUT_MS", 20000));
  if (h->cfg.tile_nnz <= 0) h->cfg.tile_nnz = env_int("B200S_TILE_NNZ", 0);
  if (h->cfg.tile_rows <= 0) h->cfg.tile_rows = env_int("B200S_TILE_ROWS", 0);
  *out = h;
  return B200S_OK;
}

void b200s_destroy(b200s_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  destroy_graphs(h->cg);
  destroy_graphs(h->bicg);
  for (GraphSet& g : h->cg_multi) destroy_graphs(g);
  if (h->hSm) cudaFreeHost(h->hSm);
  free_factors(h);
  for (int q = 0; q < static_cast<int>(h->peer_window.size()); ++q)
    if (q != h->plan.rank && h->peer_window[q]) cudaIpcCloseMemHandle(h->peer_window[q]);
  DevBuf* bufs[] = {&h->mx, &h->mr, &h->mp, &h->mq, &h->mb, &h->mS, &h->mpartials, &h->mstage,
                    &h->b, &h->r, &h->q, &h->r0, &h->s, &h->t, &h->yout, &h->invdiag,  // slices first
                    &h->rowptr, &h->colidx, &h->vals, &h->src, &h->tiles, &h->send_rows, &h->window,
                    &h->scalars, &h->partials, &h->counter,
                    &h->halo_counter, &h->history, &h->gridbar};
  for (DevBuf* b : bufs) dev_free(h, *b);
  if (h->hS) cudaFreeHost(h->hS);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int b200s_analyze_pattern(b200s_handle* h, int64_t rows, int64_t cols, int64_t nnz, const int32_t* rowptr,
                          const int32_t* colidx, const int32_t* inner_nnz, int uplo, const int64_t* row_starts) {
  if (!h) return B200S_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  h->analyzed = h->factorized = false;
  free_factors(h);
  h->precond = h->precond_factorize;
  destroy_graphs(h->cg);
  destroy_graphs(h->bicg);
  for (GraphSet& g : h->cg_multi) destroy_graphs(g);
  std::string err;
  int rc = build_plan(h->cfg, rows, cols, nnz, rowptr, colidx, inner_nnz, uplo, row_starts, h->plan, err);
  if (rc) return fail(h, rc, err);
  const Plan& p = h->plan;
  if ((rc = configure_spmv(h))) return rc;

  // ---- pattern -> device ----
  if ((rc = dev_alloc(h, h->rowptr, static_cast<size_t>(p.rows + 1) * 4 + 64))) return rc;
  if ((rc = dev_alloc(h, h->colidx, static_cast<size_t>(p.nnz) * 4 + 64))) return rc;
  if ((rc = dev_alloc(h, h->tiles, std::max<size_t>(1, p.tiles.size()) * sizeof(Tile)))) return rc;
  CK(cudaMemcpyAsync(h->rowptr.p, p.rowptr.data(), static_cast<size_t>(p.rows + 1) * 4, cudaMemcpyHostToDevice, h->stream));
  if (p.nnz) CK(cudaMemcpyAsync(h->colidx.p, p.colidx_ptr(), static_cast<size_t>(p.nnz) * 4, cudaMemcpyHostToDevice, h->stream));
  if (!p.tiles.empty()) CK(cudaMemcpyAsync(h->tiles.p, p.tiles.data(), p.tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice, h->stream));
  if (!p.src.empty()) {
    if ((rc = dev_alloc(h, h->src, p.src.size() * 4))) return rc;
    CK(cudaMemcpyAsync(h->src.p, p.src.data(), p.src.size() * 4, cudaMemcpyHostToDevice, h->stream));
  }
  if (!p.send_rows.empty()) {
    if ((rc = dev_alloc(h, h->send_rows, p.send_rows.size() * 4))) return rc;
    CK(cudaMemcpyAsync(h->send_rows.p, p.send_rows.data(), p.send_rows.size() * 4, cudaMemcpyHostToDevice, h->stream));
  }
  if ((rc = dev_alloc(h, h->scalars, sizeof(Scalars), true))) return rc;
  if ((rc = dev_alloc(h, h->partials, sizeof(double) * kMaxGrid * 4, true))) return rc;
  if ((rc = dev_alloc(h, h->counter, 64, true))) return rc;
  if ((rc = dev_alloc(h, h->halo_counter, 64, true))) return rc;
  if ((rc = dev_alloc(h, h->history, sizeof(double) * kHistoryCap, true))) return rc;
  if ((rc = dev_alloc(h, h->gridbar, 8192, true))) return rc;

  // ---- peer-visible window: 4 extended vector slots + all-reduce mailboxes + halo flags ----
  const int W = p.world;
  h->all_rows.assign(W, 0);
  h->all_ghosts.assign(W, 0);
  // one GPU: a slot must hold an operand of the product (cols entries) as well as a result (rows entries)
  int64_t mine[2] = {W == 1 ? std::max(p.rows, p.cols) : p.rows, static_cast<int64_t>(p.ghost_cols.size())};
  if (W > 1) {
    std::vector<int64_t> all(2 * W);
    if (h->cfg.allgather(h->cfg.allgather_ctx, mine, all.data(), sizeof(mine))) return fail(h, B200S_ERR_COMM, "allgather(rows, ghosts) failed");
    for (int q = 0; q < W; ++q) { h->all_rows[q] = all[2 * q]; h->all_ghosts[q] = all[2 * q + 1]; }
  } else {
    h->all_rows[0] = mine[0];
    h->all_ghosts[0] = mine[1];
  }
  h->slot_bytes = ext_slot_bytes(mine[0], mine[1]);
  h->slots_off = slots_off(mine[0]);
  h->box_off = h->slots_off + h->slot_bytes * kNumSlots;
  h->flag_off = h->box_off + sizeof(unsigned long long) * 2 * kMaxWorld * kBoxWords;
  h->halo_flag_off = h->flag_off + sizeof(unsigned) * 2 * kMaxWorld;
  const size_t back_off = (h->halo_flag_off + sizeof(unsigned) * kMaxWorld + 511) & ~size_t(255);
  const size_t pv = priv_vec_bytes(mine[0]);
  h->window_bytes = back_off + kBackVecs * pv;
  for (int q = 0; q < static_cast<int>(h->peer_window.size()); ++q)
    if (q != p.rank && h->peer_window[q]) cudaIpcCloseMemHandle(h->peer_window[q]);
  h->peer_window.assign(W, nullptr);
  {
    DevBuf* slices[] = {&h->r, &h->q, &h->invdiag, &h->b, &h->r0, &h->s, &h->t, &h->yout};
    for (DevBuf* sl : slices) dev_free(h, *sl);
    if (h->window.p) { dev_free(h, h->window); }
    if ((rc = dev_alloc(h, h->window, h->window_bytes, true))) return rc;
    char* base = static_cast<char*>(h->window.p);
    for (int i = 0; i < 8; ++i) {
      slices[i]->p = base + (i < kFrontVecs ? i * pv : back_off + (i - kFrontVecs) * pv);
      slices[i]->bytes = pv;
      slices[i]->owned = false;
    }
  }
  if ((rc = configure_l2_persistence(h))) return rc;
  CK(cudaStreamSynchronize(h->stream));
  h->peer_window[p.rank] = h->window.p;
  if (W > 1) {
    cudaIpcMemHandle_t mh;
    CK(cudaIpcGetMemHandle(&mh, h->window.p));
    std::vector<cudaIpcMemHandle_t> all(W);
    if (h->cfg.allgather(h->cfg.allgather_ctx, &mh, all.data(), sizeof(mh))) return fail(h, B200S_ERR_COMM, "allgather(ipc handles) failed");
    for (int q = 0; q < W; ++q)
      if (q != p.rank) CK(cudaIpcOpenMemHandle(&h->peer_window[q], all[q], cudaIpcMemLazyEnablePeerAccess));
    // nobody may start pushing before everyone has mapped everyone
    int64_t token = 1;
    std::vector<int64_t> tokens(W);
    if (h->cfg.allgather(h->cfg.allgather_ctx, &token, tokens.data(), sizeof(token))) return fail(h, B200S_ERR_COMM, "allgather(barrier) failed");
  }
  h->analyzed = true;
  return B200S_OK;
}

int b200s_factorize_f64(b200s_handle* h, const double* values, int precond) {
  if (!h) return B200S_ERR_INVALID;
  return factorize_impl<double>(h, values, precond);
}
int b200s_factorize_f32(b200s_handle* h, const float* values, int precond) {
  if (!h) return B200S_ERR_INVALID;
  return factorize_impl<float>(h, values, precond);
}

int b200s_spmv_f64(b200s_handle* h, const double* x, double* y) { return h ? spmv_host_impl<double>(h, x, y) : B200S_ERR_INVALID; }
int b200s_spmv_f32(b200s_handle* h, const float* x, float* y) { return h ? spmv_host_impl<float>(h, x, y) : B200S_ERR_INVALID; }
int b200s_spmv_device_f64(b200s_handle* h, const double* x, double* y, int reps, float* ms) {
  return h ? spmv_device_impl<double>(h, x, y, reps, ms) : B200S_ERR_INVALID;
}
int b200s_spmv_device_f32(b200s_handle* h, const float* x, float* y, int reps, float* ms) {
  return h ? spmv_device_impl<float>(h, x, y, reps, ms) : B200S_ERR_INVALID;
}

#define B200S_SOLVE_ENTRY(NAME, T, IMPL, BICG)                                                                      \
  int NAME(b200s_handle* h, const T* b, T* x, int use_guess, double tol, int64_t max_iters, int64_t* iters_out,      \
           double* error_out, int* info_out) {                                                                       \
    return IMPL<T>(h, BICG, b, x, use_guess, tol, max_iters, iters_out, error_out, info_out);                        \
  }
// double: an incomplete-factorization preconditioner (b200s_set_preconditioner) routes to the general loops of precond.inc
#define B200S_SOLVE_ENTRY_F64(NAME, IMPL, BICG, DEVICE)                                                              \
  int NAME(b200s_handle* h, const double* b, double* x, int use_guess, double tol, int64_t max_iters,                \
           int64_t* iters_out, double* error_out, int* info_out) {                                                   \
    if (h && h->precond == B200S_PRECOND_FACTORS)                                                                    \
      return (BICG ? bicgstab_general_solve : cg_general_solve)(h, b, x, DEVICE, use_guess, tol, max_iters,          \
                                                                iters_out, error_out, info_out);                    \
    return IMPL<double>(h, BICG, b, x, use_guess, tol, max_iters, iters_out, error_out, info_out);                   \
  }
B200S_SOLVE_ENTRY_F64(b200s_cg_solve_f64, solve_host, false, false)
B200S_SOLVE_ENTRY_F64(b200s_bicgstab_solve_f64, solve_host, true, false)
B200S_SOLVE_ENTRY(b200s_cg_solve_f32, float, solve_host, false)
B200S_SOLVE_ENTRY(b200s_bicgstab_solve_f32, float, solve_host, true)
B200S_SOLVE_ENTRY_F64(b200s_cg_solve_device_f64, solve_device, false, true)
B200S_SOLVE_ENTRY_F64(b200s_bicgstab_solve_device_f64, solve_device, true, true)
B200S_SOLVE_ENTRY(b200s_cg_solve_device_f32, float, solve_device, false)
B200S_SOLVE_ENTRY(b200s_bicgstab_solve_device_f32, float, solve_device, true)
#undef B200S_SOLVE_ENTRY
#undef B200S_SOLVE_ENTRY_F64

// ---- incomplete factorizations (host analysis: csrc/factors.cpp; device application: precond.inc) ----
#define B200S_NEW_FACTORS(CALL)                                  \
  if (!out) return B200S_ERR_INVALID;                            \
  *out = nullptr;                                                \
  b200s_factors* f = new (std::nothrow) b200s_factors();         \
  if (!f) return B200S_ERR_ALLOC;                                \
  std::string err;                                               \
  int rc;                                                        \
  try {                                                          \
    rc = CALL;                                                   \
  } catch (const std::bad_alloc&) { /* no exception crosses the C boundary */ \
    rc = B200S_ERR_ALLOC;                                        \
    err = "host allocation failed while factorizing";           \
  }                                                              \
  if (rc) {                                                      \
    g_create_error = err;                                        \
    delete f;                                                    \
    return rc;                                                   \
  }                                                              \
  *out = f;                                                      \
  return 0;

int b200s_ilut_f64(int64_t n, const int32_t* rowptr, const int32_t* colidx, const double* values, double droptol,
                   int fillfactor, const int32_t* perm, b200s_factors** out) {
  B200S_NEW_FACTORS(ilut_factorize(n, rowptr, colidx, values, droptol, fillfactor, perm, f->f, err))
}
int b200s_ichol_f64(int64_t n, const int32_t* rowptr, const int32_t* colidx, const double* values, int uplo,
                    double initial_shift, const int32_t* perm, b200s_factors** out) {
  B200S_NEW_FACTORS(ichol_factorize(n, rowptr, colidx, values, uplo, initial_shift, perm, f->f, err))
}
int b200s_factors_from_ilut_f64(int64_t n, const int32_t* lu_rowptr, const int32_t* lu_colidx, const double* lu_values,
                                const int32_t* perm, b200s_factors** out) {
  B200S_NEW_FACTORS(factors_from_ilut(n, lu_rowptr, lu_colidx, lu_values, perm, f->f, err))
}
int b200s_factors_from_ichol_f64(int64_t n, const int32_t* colptr, const int32_t* rowidx, const double* l_values,
                                 const double* scale, const int32_t* perm, b200s_factors** out) {
  B200S_NEW_FACTORS(factors_from_ichol(n, colptr, rowidx, l_values, scale, perm, f->f, err))
}
#undef B200S_NEW_FACTORS
int b200s_ordering_multicolor(int64_t n, const int32_t* rowptr, const int32_t* colidx, int32_t* perm) {
  std::string err;
  int rc;
  try {
    rc = multicolor_ordering(n, rowptr, colidx, perm, err);
  } catch (const std::bad_alloc&) {
    rc = B200S_ERR_ALLOC;
    err = "host allocation failed while ordering";
  }
  if (rc < 0) g_create_error = err;
  return rc;
}
void b200s_factors_destroy(b200s_factors* f) { delete f; }
int b200s_factors_info(const b200s_factors* f) { return f ? f->f.info : 3; }
int b200s_factors_kind(const b200s_factors* f) { return f ? f->f.kind : 0; }
int64_t b200s_factors_size(const b200s_factors* f) { return f ? f->f.n : B200S_ERR_INVALID; }
int64_t b200s_factors_nnz(const b200s_factors* f) { return f ? static_cast<int64_t>(f->f.inner.size()) : B200S_ERR_INVALID; }
int64_t b200s_factors_perm_size(const b200s_factors* f) { return f ? static_cast<int64_t>(f->f.perm.size()) : B200S_ERR_INVALID; }
int b200s_factors_get(const b200s_factors* f, int32_t* outer, int32_t* inner, double* values, double* scale, int32_t* perm) {
  if (!f) return B200S_ERR_INVALID;
  const Factors& F = f->f;
  if (outer) std::copy(F.outer.begin(), F.outer.end(), outer);
  if (inner) std::copy(F.inner.begin(), F.inner.end(), inner);
  if (values) std::copy(F.vals.begin(), F.vals.end(), values);
  if (scale) std::copy(F.scale.begin(), F.scale.end(), scale);
  if (perm) std::copy(F.perm.begin(), F.perm.end(), perm);
  return 0;
}
int b200s_factors_stage_sizes(const b200s_factors* f, int which, int64_t* nnz, int32_t* levels, int32_t* launches,
                              int32_t* unit_diag, int32_t* fused) {
  if (!f || which < 0 || which > 1) return B200S_ERR_INVALID;
  const TriStage& s = which ? f->f.second : f->f.first;
  if (nnz) *nnz = static_cast<int64_t>(s.colidx.size());
  if (levels) *levels = s.level_ptr.empty() ? 0 : static_cast<int32_t>(s.level_ptr.size()) - 1;
  if (launches) *launches = static_cast<int32_t>(s.launches.size());
  if (unit_diag) *unit_diag = s.diag.empty() ? 1 : 0;
  if (fused) *fused = s.fused ? 1 : 0;
  return 0;
}
int b200s_factors_stage(const b200s_factors* f, int which, int32_t* rowptr, int32_t* colidx, double* values, double* diag,
                        int32_t* level_ptr, int32_t* level_rows, int32_t* launches) {
  if (!f || which < 0 || which > 1) return B200S_ERR_INVALID;
  const TriStage& s = which ? f->f.second : f->f.first;
  if (rowptr) std::copy(s.rowptr.begin(), s.rowptr.end(), rowptr);
  if (colidx) std::copy(s.colidx.begin(), s.colidx.end(), colidx);
  if (values) std::copy(s.vals.begin(), s.vals.end(), values);
  if (diag) std::copy(s.diag.begin(), s.diag.end(), diag);
  if (level_ptr) std::copy(s.level_ptr.begin(), s.level_ptr.end(), level_ptr);
  if (level_rows) std::copy(s.level_rows.begin(), s.level_rows.end(), level_rows);
  if (launches)
    for (size_t i = 0; i < s.launches.size(); ++i) {
      launches[3 * i] = s.launches[i].level_begin;
      launches[3 * i + 1] = s.launches[i].level_end;
      launches[3 * i + 2] = s.launches[i].rows;
    }
  return 0;
}
int b200s_factors_permscale(const b200s_factors* f, int32_t* pre_gather, double* pre_scale, int32_t* post_gather,
                            double* post_scale, int32_t* present) {
  if (!f) return B200S_ERR_INVALID;
  const Factors& F = f->f;
  if (pre_gather) std::copy(F.pre_gather.begin(), F.pre_gather.end(), pre_gather);
  if (pre_scale) std::copy(F.pre_scale.begin(), F.pre_scale.end(), pre_scale);
  if (post_gather) std::copy(F.post_gather.begin(), F.post_gather.end(), post_gather);
  if (post_scale) std::copy(F.post_scale.begin(), F.post_scale.end(), post_scale);
  if (present) {
    present[0] = !F.pre_gather.empty();
    present[1] = !F.pre_scale.empty();
    present[2] = !F.post_gather.empty();
    present[3] = !F.post_scale.empty();
  }
  return 0;
}
int b200s_set_preconditioner(b200s_handle* h, const b200s_factors* f) {
  if (!h) return B200S_ERR_INVALID;
  try {
    return set_factors(h, f ? &f->f : nullptr);
  } catch (const std::bad_alloc&) {  // no exception crosses the C boundary
    return fail(h, B200S_ERR_ALLOC, "set_preconditioner: host allocation failed");
  }
}
int b200s_precond_apply_f64(b200s_handle* h, const double* r, double* z) { return precond_apply_host(h, r, z); }

int b200s_lscg_solve_f64(b200s_handle* hA, b200s_handle* hAt, const double* b, double* x, int use_guess, double tol,
                         int64_t max_iters, int precond, int colmajor_precond, int64_t* iters_out, double* error_out,
                         int* info_out) {
  return lscg_solve(hA, hAt, b, x, use_guess, tol, max_iters, precond, colmajor_precond, iters_out, error_out, info_out);
}
int b200s_minres_solve_f64(b200s_handle* h, const double* b, double* x, int use_guess, double tol, int64_t max_iters,
                           int64_t* iters_out, double* error_out, int* info_out) {
  return minres_solve(h, b, x, use_guess, tol, max_iters, iters_out, error_out, info_out);
}
int b200s_gmres_solve_f64(b200s_handle* h, const double* b, double* x, int use_guess, double tol, int64_t max_iters,
                          int64_t restart, int64_t* iters_out, double* error_out, int* info_out) {
  return gmres_solve(h, b, x, use_guess, tol, max_iters, restart, iters_out, error_out, info_out);
}

int b200s_multi_rhs_batch(b200s_handle* h) {
  return (h && multi_supported(h)) ? std::max(2, std::min(kMultiMax, env_int("B200S_MULTI_K", 4))) : 0;
}

int b200s_cg_solve_multi_device_f64(b200s_handle* h, int64_t ncols, const double* B_dev, int64_t ldb, double* X_dev,
                                    int64_t ldx, int use_guess, double tol, int64_t max_iters, int64_t* iters_out,
                                    double* error_out, int* info_out) {
  if (!h) return B200S_ERR_INVALID;
  if (!h->factorized) return fail(h, B200S_ERR_INVALID, "solve: call analyze_pattern + factorize first");
  if (h->scalar_bytes != 8) return fail(h, B200S_ERR_INVALID, "solve_multi: the matrix was factorized in float");
  if (ncols < 0 || (ncols > 0 && (!B_dev || !X_dev)) || ldb < h->plan.rows || ldx < h->plan.rows)
    return fail(h, B200S_ERR_INVALID, "solve_multi: bad argument");
  CK(cudaSetDevice(h->device));
  return solve_multi_device(h, ncols, B_dev, ldb, X_dev, ldx, use_guess, tol, max_iters, iters_out, error_out, info_out,
                            nullptr);
}

int b200s_cg_solve_multi_f64(b200s_handle* h, int64_t ncols, const double* B, int64_t ldb, double* X, int64_t ldx,
                             int use_guess, double tol, int64_t max_iters, int64_t* iters_out, double* error_out,
                             int* info_out) {
  if (!h) return B200S_ERR_INVALID;
  if (!h->factorized) return fail(h, B200S_ERR_INVALID, "solve: call analyze_pattern + factorize first");
  if (h->scalar_bytes != 8) return fail(h, B200S_ERR_INVALID, "solve_multi: the matrix was factorized in float");
  const int64_t n = h->plan.rows;
  if (ncols < 0 || (ncols > 0 && (!B || !X)) || ldb < n || ldx < n) return fail(h, B200S_ERR_INVALID, "solve_multi: bad argument");
  if (ncols == 0) return 0;
  CK(cudaSetDevice(h->device));
  int rc;
  // device staging: [B | X], column-major with leading dimension n
  if ((rc = dev_alloc(h, h->mstage, static_cast<size_t>(n) * ncols * 16))) return rc;
  double* Bd = h->mstage.as<double>();
  double* Xd = Bd + static_cast<size_t>(n) * ncols;
  cudaEvent_t e0 = h->ev0, e1 = h->ev1;
  CK(cudaEventRecord(e0, h->stream));
  CK(cudaMemcpy2DAsync(Bd, n * 8, B, ldb * 8, n * 8, ncols, cudaMemcpyHostToDevice, h->stream));
  if (use_guess) CK(cudaMemcpy2DAsync(Xd, n * 8, X, ldx * 8, n * 8, ncols, cudaMemcpyHostToDevice, h->stream));
  CK(cudaEventRecord(e1, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double h2d = ms;
  if ((rc = solve_multi_device(h, ncols, Bd, n, Xd, n, use_guess, tol, max_iters, iters_out, error_out, info_out, nullptr)))
    return rc;
  const double solve_ms = h->last_solve_ms;
  CK(cudaEventRecord(e0, h->stream));
  CK(cudaMemcpy2DAsync(X, ldx * 8, Xd, n * 8, n * 8, ncols, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaEventRecord(e1, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  h->last_h2d_ms = h2d;
  h->last_d2h_ms = ms;
  h->last_solve_ms = solve_ms;
  return 0;
}

int b200s_get_stats(b200s_handle* h, b200s_stats* out) {
  if (!h || !out) return B200S_ERR_INVALID;
  b200s_stats st;
  std::memset(&st, 0, sizeof(st));
  fill_tile_stats(h->plan, &st);
  st.device = h->device;
  st.spmv_grid = h->spmv_grid;
  st.spmv_block = kSpmvThreads;
  st.spmv_smem_bytes = h->spmv_smem;
  st.spmv_stages = h->spmv_stages;
  st.vec_grid = h->vec_grid;
  st.vec_block = kVecThreads;
  st.loop_mode = h->loop_mode;
  st.evict_first = h->evict_first;
  st.l2_persist = h->l2_persist;
  st.sm_count = h->sm_count;
  st.last_solve_ms = h->last_solve_ms;
  st.last_h2d_ms = h->last_h2d_ms;
  st.last_d2h_ms = h->last_d2h_ms;
  st.last_kernel_launches = h->last_launches;
  st.last_iterations = h->last_iterations;
  st.last_spmv_count = h->last_spmv;
  st.last_restarts = h->last_restarts;
  st.last_nonfinite = h->last_nonfinite;
  st.last_comm_error = h->last_comm_error;
  st.device_bytes = static_cast<int64_t>(h->device_bytes);
  int32_t want = out->struct_size > 0 ? out->struct_size : static_cast<int32_t>(sizeof(b200s_stats));
  st.struct_size = std::min<int32_t>(want, sizeof(b200s_stats));
  std::memcpy(out, &st, st.struct_size);
  return B200S_OK;
}

int b200s_get_invdiag_f64(b200s_handle* h, double* invdiag) {
  if (!h || !invdiag) return B200S_ERR_INVALID;
  if (!h->factorized || h->scalar_bytes != 8) return fail(h, B200S_ERR_INVALID, "get_invdiag: factorize_f64 first");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(invdiag, h->invdiag.p, static_cast<size_t>(h->plan.rows) * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return B200S_OK;
}

int b200s_get_timeline(b200s_handle* h, double* out, int cap) {
  // out[0..11] = microseconds per epilogue kind (kernel + launch gap up to that reduction), out[12..23] = counts,
  // out[24] = all-reduce us, out[25] = halo wait us (CTA 0), out[26] = first->last epilogue us
  if (!h || !out || cap < 27 || !h->hS) return B200S_ERR_INVALID;
  const Scalars& S = *h->hS;
  for (int i = 0; i < 12; ++i) { out[i] = S.t_phase[i] * 1e-3; out[12 + i] = static_cast<double>(S.n_phase[i]); }
  out[24] = S.t_allreduce * 1e-3;
  out[25] = S.t_halo_wait * 1e-3;
  out[26] = (S.t_end - S.t_first) * 1e-3;
  return B200S_OK;
}

int64_t b200s_get_residual_history(b200s_handle* h, double* rr, int64_t cap) {
  if (!h || !rr || cap < 0) return B200S_ERR_INVALID;
  if (!h->hS) return 0;
  int64_t n = std::min<int64_t>(cap, std::min<int64_t>(h->hS->hist_len, kHistoryCap));
  if (n > 0) {
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(rr, h->history.p, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return n;
}

int64_t b200s_plan_probe(const b200s_config* cfg, int64_t rows, int64_t cols, int64_t nnz, const int32_t* rowptr,
                         const int32_t* colidx, const int64_t* row_starts, int32_t* local_colidx,
                         int64_t* ghost_cols, int64_t ghost_cap, int32_t* send_rows, int64_t send_cap,
                         int64_t* send_counts, int64_t* recv_counts, b200s_stats* tile_stats) {
  b200s_config c;
  std::memset(&c, 0, sizeof(c));
  c.world = 1;
  if (cfg) std::memcpy(&c, cfg, std::min<size_t>(sizeof(c), cfg->struct_size > 0 ? cfg->struct_size : sizeof(c)));
  if (c.world <= 0) c.world = 1;
  Plan p;
  std::string err;
  int rc = build_plan(c, rows, cols, nnz, rowptr, colidx, nullptr, B200S_BOTH, row_starts, p, err);
  if (rc) {
    g_create_error = err;
    return rc;
  }
  if (local_colidx && p.nnz) std::memcpy(local_colidx, p.colidx_ptr(), static_cast<size_t>(p.nnz) * 4);
  if (ghost_cols)
    for (int64_t g = 0; g < std::min<int64_t>(ghost_cap, p.ghost_cols.size()); ++g) ghost_cols[g] = p.ghost_cols[g];
  if (send_rows)
    for (int64_t k = 0; k < std::min<int64_t>(send_cap, p.send_rows.size()); ++k) send_rows[k] = p.send_rows[k];
  for (int q = 0; q < p.world; ++q) {
    if (send_counts) send_counts[q] = p.send_counts[q];
    if (recv_counts) recv_counts[q] = p.recv_counts[q];
  }
  if (tile_stats) {
    int32_t want = tile_stats->struct_size > 0 ? tile_stats->struct_size : static_cast<int32_t>(sizeof(b200s_stats));
    b200s_stats st;
    std::memset(&st, 0, sizeof(st));
    fill_tile_stats(p, &st);
    st.struct_size = std::min<int32_t>(want, sizeof(b200s_stats));
    std::memcpy(tile_stats, &st, st.struct_size);
  }
  return static_cast<int64_t>(p.ghost_cols.size());
}

int64_t b200s_plan_probe_csr(int64_t rows, int64_t nnz, const int32_t* rowptr, const int32_t* colidx,
                             const int32_t* inner_nnz, int uplo, int32_t* out_rowptr, int32_t* out_colidx,
                             int32_t* out_src, int64_t cap) {
  b200s_config c;
  std::memset(&c, 0, sizeof(c));
  c.world = 1;
  Plan p;
  std::string err;
  int rc = build_plan(c, rows, rows, nnz, rowptr, colidx, inner_nnz, uplo, nullptr, p, err);
  if (rc) {
    g_create_error = err;
    return rc;
  }
  if (out_rowptr) std::memcpy(out_rowptr, p.rowptr.data(), sizeof(int32_t) * static_cast<size_t>(rows + 1));
  const int64_t n = std::min<int64_t>(cap, p.nnz);
  if (out_colidx && n > 0) std::memcpy(out_colidx, p.colidx_ptr(), sizeof(int32_t) * static_cast<size_t>(n));
  if (out_src)
    for (int64_t k = 0; k < n; ++k) out_src[k] = p.src.empty() ? static_cast<int32_t>(k) : p.src[k];
  return p.nnz;
}

int64_t b200s_plan_probe_selfadjoint(const b200s_config* cfg, int64_t rows, int64_t cols, int64_t nnz,
                                     const int32_t* rowptr, const int32_t* colidx, const int32_t* inner_nnz, int uplo,
                                     const int64_t* row_starts, const double* values, int32_t* out_rowptr,
                                     int64_t* out_cols, double* out_values, int64_t cap) {
  b200s_config c;
  std::memset(&c, 0, sizeof(c));
  c.world = 1;
  if (cfg) std::memcpy(&c, cfg, std::min<size_t>(sizeof(c), cfg->struct_size > 0 ? cfg->struct_size : sizeof(c)));
  if (c.world <= 0) c.world = 1;
  Plan p;
  std::string err;
  int rc = build_plan(c, rows, cols, nnz, rowptr, colidx, inner_nnz, uplo, row_starts, p, err);
  std::vector<unsigned char> imports;
  if (!rc && values) rc = exchange_mirror_values(c, p, values, sizeof(double), imports, err);
  if (rc) {
    g_create_error = err;
    return rc;
  }
  if (out_rowptr) std::memcpy(out_rowptr, p.rowptr.data(), sizeof(int32_t) * static_cast<size_t>(rows + 1));
  const int64_t n = std::min<int64_t>(cap, p.nnz);
  const int32_t* lc = p.colidx_ptr();
  const double* imp = reinterpret_cast<const double*>(imports.data());
  for (int64_t k = 0; k < n; ++k) {
    if (out_cols) out_cols[k] = (p.world == 1 || lc[k] < p.rows) ? p.row0 + lc[k] : p.ghost_cols[lc[k] - p.rows];
    if (out_values && values) {
      const int64_t s = p.src.empty() ? k : p.src[k];
      out_values[k] = s < p.input_nnz ? values[s] : imp[s - p.input_nnz];
    }
  }
  return p.nnz;
}

int64_t b200s_plan_probe_span(int64_t rows, int64_t nnz, const int32_t* rowptr, const int32_t* colidx,
                              const int32_t* inner_nnz, int uplo) {
  b200s_config c;
  std::memset(&c, 0, sizeof(c));
  c.world = 1;
  Plan p;
  std::string err;
  int rc = build_plan(c, rows, rows, nnz, rowptr, colidx, inner_nnz, uplo, nullptr, p, err);
  if (rc) {
    g_create_error = err;
    return rc;
  }
  return p.input_nnz;
}

}  // extern "C"

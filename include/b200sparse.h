/* b200sparse.h -- C ABI of the B200-native sparse iterative-solve hot path.
 *
 * This library replaces, on one or more NVIDIA B200 GPUs (sm_100a), exactly one path of Eigen
 * (paths relative to the reference tree):
 *
 *   y = A*x, CSR                    Eigen/src/SparseCore/SparseDenseProduct.h:26-72
 *   ConjugateGradient loop          Eigen/src/IterativeLinearSolvers/ConjugateGradient.h:26-91, :197-221
 *   BiCGSTAB loop                   Eigen/src/IterativeLinearSolvers/BiCGSTAB.h:28-107, :193-204
 *   Jacobi / identity precond.      Eigen/src/IterativeLinearSolvers/BasicPreconditioners.h:64-101, :200-222
 *   solver state and setters        Eigen/src/IterativeLinearSolvers/IterativeSolverBase.h:196-330, :399-413
 *
 * Eigen has no FFI; the reference-side binding is the CRTP solver concept (IterativeSolverBase<Derived>), in the
 * manner of Eigen/src/KLUSupport/KLUSupport.h:60-115.  include/b200/IterativeSolvers.h is that binding (C++,
 * header-only, needs Eigen); INTEGRATION.md shows it.  Python tests and bench.py bind the same symbols via ctypes.
 *
 * Conventions
 *   - every function returns 0 on success or a negative b200s_status; b200s_last_error() gives the text.
 *     No exceptions cross the boundary.  A handle is not re-entrant; distinct handles are independent.
 *   - host arrays are BORROWED for the duration of the call and copied to the device; nothing is retained.
 *   - "_device" variants take device pointers (valid on the handle's device) and never touch host memory.
 *   - there is NO CPU fallback: without a usable sm_100 device b200s_create fails with B200S_ERR_NO_DEVICE.
 *   - indices are int32 (Eigen's default StorageIndex, SparseMatrix.h:36); sizes are int64.
 *   - uplo uses Eigen's bit values (Core/util/Constants.h): 1 = Lower, 2 = Upper, 3 = Lower|Upper.
 *   - info uses Eigen's ComputationInfo (Core/util/Constants.h:430-440): 0 Success, 1 NumericalIssue,
 *     2 NoConvergence, 3 InvalidInput.
 */
#ifndef B200SPARSE_H
#define B200SPARSE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200S_VERSION 200

typedef struct b200s_handle b200s_handle;

typedef enum {
  B200S_OK = 0,
  B200S_ERR_INVALID = -1,   /* bad argument / call order (Eigen: eigen_assert or InvalidInput) */
  B200S_ERR_NO_DEVICE = -2, /* no CUDA device of compute capability 10.x: the product has no CPU path */
  B200S_ERR_CUDA = -3,      /* a CUDA runtime call failed; text in b200s_last_error */
  B200S_ERR_ALLOC = -4,
  B200S_ERR_COMM = -5,      /* multi-GPU bootstrap failed, a peer failed before a launch, or a device-side wait expired */
  B200S_ERR_UNSUPPORTED = -6
} b200s_status;

enum { B200S_LOWER = 1, B200S_UPPER = 2, B200S_BOTH = 3 };
enum { B200S_PRECOND_IDENTITY = 0, B200S_PRECOND_JACOBI = 1,
       B200S_PRECOND_FACTORS = 2 /* an incomplete factorization given to b200s_set_preconditioner */ };
enum { B200S_FACTORS_ILUT = 1, B200S_FACTORS_ICHOL = 2 };
enum { B200S_SPMV_AUTO = 0, B200S_SPMV_STAGED = 1, B200S_SPMV_DIRECT = 2 };
enum {
  B200S_LOOP_AUTO = 0,
  B200S_LOOP_WHILE_GRAPH = 1,   /* one graph launch; WHILE node iterates on a device-side condition */
  B200S_LOOP_CHUNKED_GRAPH = 2, /* graphs of chunk_iters unrolled iterations; host polls the stop flag per chunk */
  B200S_LOOP_STREAM = 3,        /* plain stream launches, host polls per iteration (debug / profiling) */
  B200S_LOOP_PERSISTENT = 4     /* CG only: the whole loop in one cooperative kernel (grid barriers); BiCGSTAB uses WHILE */
};

/* Host-provided all-gather used ONLY while building the multi-GPU plan (setup, never in the iteration):
 * every rank contributes `bytes` bytes from `send`; `recv` receives world*bytes, ordered by rank.  Return 0 on
 * success.  Python binds it to torch.distributed.all_gather, C++ hosts to MPI_Allgather. */
typedef int (*b200s_allgather_fn)(void* ctx, const void* send, void* recv, size_t bytes);

typedef struct {
  int32_t struct_size;  /* = sizeof(b200s_config); lets the struct grow */
  int32_t device;       /* CUDA device ordinal; -1 = current device */
  int32_t rank;         /* this process owns row block `rank` of `world` (one process per GPU) */
  int32_t world;        /* 1 = single GPU */
  int32_t spmv_impl;    /* B200S_SPMV_* */
  int32_t loop_mode;    /* B200S_LOOP_* */
  int32_t chunk_iters;  /* iterations per graph launch in CHUNKED mode (0 = default 32) */
  int32_t tile_nnz;     /* staged SpMV: max non-zeros per shared-memory tile (0 = default) */
  int32_t tile_rows;    /* staged SpMV: max rows per tile (0 = default) */
  int32_t reserved0;
  b200s_allgather_fn allgather; /* required when world > 1 */
  void* allgather_ctx;
} b200s_config;

typedef struct {
  int32_t struct_size;
  int32_t world, rank, device;
  int64_t rows, cols, nnz;          /* local block after compression / symmetric expansion */
  int64_t ghosts;                   /* halo entries received per SpMV */
  int64_t halo_send;                /* halo entries sent per SpMV */
  int32_t tiles, tiles_boundary;    /* staged-SpMV tiles (total / touching ghosts) */
  int32_t tiles_by_lanes[6];        /* tiles using 1,2,4,8,16,32 lanes per row */
  int32_t tiles_stream, tiles_long; /* two-phase (CSR-stream) tiles, rows longer than a tile */
  int32_t spmv_grid, spmv_block, spmv_smem_bytes, spmv_stages;
  int32_t vec_grid, vec_block;
  int32_t loop_mode;                /* resolved */
  int32_t evict_first;              /* matrix stream uses the L2 evict-first policy (auto: vectors fit L2, matrix does not) */
  int32_t sm_count;
  double last_solve_ms;             /* device time of the last solve (CUDA events around the graph launch) */
  double last_h2d_ms, last_d2h_ms;  /* host<->device copies of the last host-buffer call */
  int64_t last_kernel_launches;     /* kernels of this library launched by the last solve / spmv call */
  int64_t last_iterations;
  int64_t last_spmv_count;          /* SpMV launches inside the last solve */
  int64_t device_bytes;             /* device memory held by the handle */
  int64_t last_restarts;            /* BiCGSTAB: re-orthogonalisation restarts taken by the last solve (BiCGSTAB.h:72-81) */
  int32_t last_nonfinite;           /* CG: the last solve stopped on a non-finite residual norm (outputs as the reference's) */
  int32_t last_comm_error;          /* a bounded device-side wait on a peer expired during the last call */
  int32_t l2_persist;               /* the CG working set (r, Ap, D^-1, x, p) is marked persisting in L2 */
  int32_t reserved1;
} b200s_stats;

/* ---- lifetime ------------------------------------------------------------------------------------------------ */
int b200s_version(void);
int b200s_device_count(void); /* number of CUDA devices with compute capability 10.x (0 = product unusable) */
int b200s_create(const b200s_config* cfg /* may be NULL: defaults, device = current */, b200s_handle** out);
void b200s_destroy(b200s_handle* h);
const char* b200s_last_error(const b200s_handle* h /* NULL: error of the last failed b200s_create */);

/* ---- setup: IterativeSolverBase::analyzePattern / factorize / compute (IterativeSolverBase.h:196-247) ----------
 * analyze_pattern takes THIS RANK's row block [row_starts[rank], row_starts[rank+1]) of a square `cols` x `cols`
 * matrix in CSR with GLOBAL column indices (world == 1: the whole matrix, row_starts may be NULL).
 *   nnz       : number of slots in colidx / values, i.e. at least one past the last slot rowptr references
 *               (rowptr[rows] for compressed storage, also when rowptr[0] > 0 as for a Map of an inner panel).
 *   inner_nnz : NULL for a compressed matrix, else Eigen's innerNonZeroPtr (SparseMatrix.h:176-183): row i holds
 *               entries [rowptr[i], rowptr[i]+inner_nnz[i]).
 *   uplo      : which stored triangle(s) define the operator.  LOWER / UPPER read one triangle and imply its mirror
 *               image, as ConjugateGradient<_,Lower> does through selfadjointView (ConjugateGradient.h:202-213); the
 *               device matrix is then the expanded full CSR.  With world > 1 the mirror image of an entry (i, c) belongs to
 *               the rank that owns row c: analyze_pattern exchanges those patterns and factorize their values through
 *               the config's allgather (setup only), so every rank passes just the triangle of its own rows.
 * It builds the partition / halo plan, the SpMV tiles with their per-tile row-binning, and uploads the pattern.
 * factorize uploads the values (same order as the pattern given to analyze_pattern) and builds the
 * preconditioner: invdiag[j] = 1/A_jj if stored and non-zero else 1 (BasicPreconditioners.h:64-79). */
int b200s_analyze_pattern(b200s_handle* h, int64_t rows, int64_t cols, int64_t nnz, const int32_t* rowptr,
                          const int32_t* colidx, const int32_t* inner_nnz, int uplo, const int64_t* row_starts);
int b200s_factorize_f64(b200s_handle* h, const double* values, int precond);
int b200s_factorize_f32(b200s_handle* h, const float* values, int precond);

/* ---- SpMV: dst = A*x (SparseDenseProduct.h:26-72 + ProductEvaluators.h:348-349) ----------------------------------
 * Host variant: x has `cols` entries when world == 1, else this rank's `rows` owned entries (the halo is exchanged
 * on the device); y receives this rank's `rows` entries.
 * Device variant: same, device pointers; runs `reps` back-to-back products (benchmark path, no PCIe) and returns
 * the average device time per product in *ms_avg (CUDA events on the library's stream). */
int b200s_spmv_f64(b200s_handle* h, const double* x, double* y);
int b200s_spmv_f32(b200s_handle* h, const float* x, float* y);
int b200s_spmv_device_f64(b200s_handle* h, const double* x_dev, double* y_dev, int reps, float* ms_avg);
int b200s_spmv_device_f32(b200s_handle* h, const float* x_dev, float* y_dev, int reps, float* ms_avg);

/* ---- solves: Derived::_solve_vector_with_guess_impl (ConjugateGradient.h:197-221, BiCGSTAB.h:193-204) ------------
 * b, x: this rank's `rows` entries.  use_guess == 0 -> x is zeroed first (IterativeSolverBase.h:399-404), else x
 * holds the initial guess (solveWithGuess, :316-323).  tol < 0 -> machine epsilon (:413); max_iters < 0 -> 2*cols
 * (:281-284).  Outputs follow the reference exactly, including: CG counts completed iterations only
 * (ConjugateGradient.h:77-79,87); ||b|| == 0 gives x = 0 with iters = 0 / error = 0 for CG but iters = max_iters /
 * error = tol for BiCGSTAB (BiCGSTAB.h:47-51); info = error <= tol ? Success : NoConvergence.  A non-finite residual
 * norm: the reference's CG loop spins on NaNs until max_iters and returns NoConvergence, iters = max_iters,
 * error = NaN, x = NaN; this library stops at once and REPORTS those same outputs (b200s_stats.last_nonfinite says
 * that it happened); BiCGSTAB leaves its loop on a NaN norm exactly as the reference does (BiCGSTAB.h:67).
 * The whole iteration runs on the device (CUDA graph, on-device convergence test); the host blocks until it is
 * finished.  The _f32 variants are the float instantiations (ConjugateGradient<SparseMatrix<float>>, ...): vectors,
 * elementwise arithmetic and the scalar recurrences in float, dot products accumulated in double and rounded once;
 * tol < 0 -> FLT_EPSILON.  They need factorize_f32.
 * Row-partitioned runs (world > 1): every rank must make the same call; ranks exchange their status through the
 * config's allgather before anything that waits on a peer is launched, and every device-side wait on a peer is
 * bounded (B200S_COMM_TIMEOUT_MS, default 20000): a dead or diverged rank yields B200S_ERR_COMM, not a hang. */
int b200s_cg_solve_f64(b200s_handle* h, const double* b, double* x, int use_guess, double tol, int64_t max_iters,
                       int64_t* iters_out, double* error_out, int* info_out);
int b200s_bicgstab_solve_f64(b200s_handle* h, const double* b, double* x, int use_guess, double tol,
                             int64_t max_iters, int64_t* iters_out, double* error_out, int* info_out);
int b200s_cg_solve_device_f64(b200s_handle* h, const double* b_dev, double* x_dev, int use_guess, double tol,
                              int64_t max_iters, int64_t* iters_out, double* error_out, int* info_out);
int b200s_bicgstab_solve_device_f64(b200s_handle* h, const double* b_dev, double* x_dev, int use_guess, double tol,
                                    int64_t max_iters, int64_t* iters_out, double* error_out, int* info_out);
int b200s_cg_solve_f32(b200s_handle* h, const float* b, float* x, int use_guess, double tol, int64_t max_iters,
                       int64_t* iters_out, double* error_out, int* info_out);
int b200s_bicgstab_solve_f32(b200s_handle* h, const float* b, float* x, int use_guess, double tol,
                             int64_t max_iters, int64_t* iters_out, double* error_out, int* info_out);
int b200s_cg_solve_device_f32(b200s_handle* h, const float* b_dev, float* x_dev, int use_guess, double tol,
                              int64_t max_iters, int64_t* iters_out, double* error_out, int* info_out);
int b200s_bicgstab_solve_device_f32(b200s_handle* h, const float* b_dev, float* x_dev, int use_guess, double tol,
                                    int64_t max_iters, int64_t* iters_out, double* error_out, int* info_out);

/* ---- multi-column right-hand sides: IterativeSolverBase::_solve_with_guess_impl, the loop of :366-389 --------------
 * B and X are column-major (Eigen's dense default) with leading dimensions ldb, ldx >= rows; column k of X receives
 * the solution for column k of B, and iters_out / error_out / info_out (ncols entries each, optional) what
 * iterations() / error() / info() would report after solving that column alone.  The reference solves the columns one
 * after the other (it streams the matrix once per column and iteration); here 4 columns at a time share one stream of the
 * matrix per iteration (SpMM with interleaved operands, one scalar recurrence per column, a converged column is
 * frozen while the others continue) and every column's result is BIT-IDENTICAL to its single-column solve.  Handles
 * the batched kernels do not cover (row-partitioned, float, matrices with very irregular rows) solve column by column
 * inside the same call; b200s_multi_rhs_batch says which (0 = column by column, else the batch width). */
int b200s_multi_rhs_batch(b200s_handle* h);
int b200s_cg_solve_multi_f64(b200s_handle* h, int64_t ncols, const double* B, int64_t ldb, double* X, int64_t ldx,
                             int use_guess, double tol, int64_t max_iters, int64_t* iters_out, double* error_out,
                             int* info_out);
int b200s_cg_solve_multi_device_f64(b200s_handle* h, int64_t ncols, const double* B_dev, int64_t ldb, double* X_dev,
                                    int64_t ldx, int use_guess, double tol, int64_t max_iters, int64_t* iters_out,
                                    double* error_out, int* info_out);

/* ---- further solvers on the same device primitives (SpMV, fused vector updates, deterministic dots) -------------------
 * One GPU, double.  The loops are the reference's, statement by statement; vectors stay on the device, the scalar
 * bookkeeping (Givens / Householder coefficients) runs on the host between kernels.
 *   lscg   : LeastSquaresConjugateGradient (Eigen/src/IterativeLinearSolvers/LeastSquareConjugateGradient.h:26-93,
 *            :190-208): min |A x - b| for a rows x cols matrix.  hA holds A, hAt holds A^T (analyze_pattern +
 *            factorize_f64 on the transposed CSR arrays; any precond value).  b has rows entries, x cols entries.
 *            precond = 1: LeastSquareDiagonalPreconditioner (BasicPreconditioners.h:127-191), inverse squared column
 *            norms; colmajor_precond selects which of the reference's two branches defines an empty column
 *            (0: row-major A -> 0, 1: column-major A -> 1).
 *   minres : MINRES (unsupported/Eigen/src/IterativeSolvers/MINRES.h:29-140, :236-262) for a self-adjoint operator
 *            (uplo of analyze_pattern as for CG); preconditioner = the one given to factorize.
 *   gmres  : restarted GMRES with Householder Arnoldi (unsupported/Eigen/src/IterativeSolvers/GMRES.h:55-212,
 *            :317-325); restart <= 0 -> 30.  info = NumericalIssue only if the reference's routine would return false.
 * tol < 0 -> machine epsilon, max_iters < 0 -> 2*cols, as for the other solvers. */
int b200s_lscg_solve_f64(b200s_handle* hA, b200s_handle* hAt, const double* b, double* x, int use_guess, double tol,
                         int64_t max_iters, int precond, int colmajor_precond, int64_t* iters_out, double* error_out,
                         int* info_out);
int b200s_minres_solve_f64(b200s_handle* h, const double* b, double* x, int use_guess, double tol, int64_t max_iters,
                           int64_t* iters_out, double* error_out, int* info_out);
int b200s_gmres_solve_f64(b200s_handle* h, const double* b, double* x, int use_guess, double tol, int64_t max_iters,
                          int64_t restart, int64_t* iters_out, double* error_out, int* info_out);

/* ---- incomplete factorizations as preconditioners --------------------------------------------------
 * IncompleteLUT (Eigen/src/IterativeLinearSolvers/IncompleteLUT.h:98-446) and IncompleteCholesky
 * (IncompleteCholesky.h:40-388) for the solvers above.  What runs EVERY ITERATION -- z = M^-1 r: two sparse triangular
 * solves between permutations / scalings (IncompleteLUT.h:171-176, IncompleteCholesky.h:149-157, TriangularSolver.h:26-134)
 * -- runs on the GPU: the rows of each solve are grouped into dependency levels, one thread sums one row in the
 * reference's order with the reference's roundings, so z has the bits of the reference as g++ -O3 builds it for x86-64 with FMA.
 * The factorization is sequential by construction (row ii needs the finished rows < ii) and is setup work: it is done
 * once per matrix on the host, either by the restatements b200s_ilut_f64 / b200s_ichol_f64 (same factors as the
 * reference, entry for entry, for the same permutation) or by the caller (b200s_factors_from_*: the C++ binding hands
 * over what an Eigen::IncompleteLUT / IncompleteCholesky object computed).  These host functions are GPU-free.
 *   perm : the fill-reducing permutation as the reference stores it -- IncompleteLUT::m_P.indices() resp.
 *          IncompleteCholesky::m_perm.indices() -- or NULL for the natural ordering.  The reference obtains it from
 *          AMDOrdering (an ordering heuristic, outside this path); any permutation is valid.
 *   b200s_ilut_f64  : droptol < 0 -> 1e-12, fillfactor <= 0 -> 10 (the reference's defaults, IncompleteLUT.h:117-118).
 *   b200s_ichol_f64 : uplo = the triangle of the CSR input that is read (LOWER or UPPER); initial_shift < 0 -> 1e-3.
 *   b200s_factors_info : Eigen's ComputationInfo of the factorization (0 Success, 1 NumericalIssue).
 * b200s_set_preconditioner (one GPU, double, after factorize) uploads the factors; from then on cg / bicgstab / minres /
 * gmres solves of the handle precondition with them (CG and BiCGSTAB then run ConjugateGradient.h:26-91 / BiCGSTAB.h:28-107
 * statement by statement on the device primitives, scalar recurrences on the host, because a triangular solve cannot be
 * fused into their vector passes).  NULL returns to the preconditioner factorize built; factorize / analyze_pattern drop
 * the factors (the matrix changed).  b200s_precond_apply_f64: z = M^-1 r for host vectors. */
typedef struct b200s_factors b200s_factors;
int b200s_ilut_f64(int64_t n, const int32_t* rowptr, const int32_t* colidx, const double* values, double droptol,
                   int fillfactor, const int32_t* perm, b200s_factors** out);
int b200s_ichol_f64(int64_t n, const int32_t* rowptr, const int32_t* colidx, const double* values, int uplo,
                    double initial_shift, const int32_t* perm, b200s_factors** out);
/* An ordering for the GPU rather than for fill: greedy multi-colouring of the pattern of A + A^T, rows sorted by colour
 * (host, GPU-free).  Rows of one colour are not coupled, so the triangular solves of a zero-fill factor have one
 * dependency level per colour (red-black on the 5/7-point stencils: 2 wide levels instead of ~3n narrow ones) at the
 * price of some extra iterations.  perm (n entries) is in the convention of the functions above; returns the number of
 * colours or a negative status. */
int b200s_ordering_multicolor(int64_t n, const int32_t* rowptr, const int32_t* colidx, int32_t* perm);
/* lu_*: IncompleteLUT::m_lu (row-major; per row: lower part, diagonal, upper part, each in any order) */
int b200s_factors_from_ilut_f64(int64_t n, const int32_t* lu_rowptr, const int32_t* lu_colidx, const double* lu_values,
                                const int32_t* perm, b200s_factors** out);
/* colptr / rowidx / l_values: IncompleteCholesky::matrixL() (column-major lower); scale: scalingS() or NULL */
int b200s_factors_from_ichol_f64(int64_t n, const int32_t* colptr, const int32_t* rowidx, const double* l_values,
                                 const double* scale, const int32_t* perm, b200s_factors** out);
void b200s_factors_destroy(b200s_factors* f);
int b200s_factors_info(const b200s_factors* f);
int b200s_factors_kind(const b200s_factors* f);          /* B200S_FACTORS_* */
int64_t b200s_factors_size(const b200s_factors* f);      /* n */
int64_t b200s_factors_nnz(const b200s_factors* f);       /* stored entries of m_lu / m_L */
int64_t b200s_factors_perm_size(const b200s_factors* f); /* n, or 0 for IncompleteCholesky with the natural ordering */
/* The factor as the reference stores it: outer (n+1), inner / values (nnz), scale (n, ICHOL), perm (perm_size). */
int b200s_factors_get(const b200s_factors* f, int32_t* outer, int32_t* inner, double* values, double* scale, int32_t* perm);
int b200s_set_preconditioner(b200s_handle* h, const b200s_factors* f);
int b200s_precond_apply_f64(b200s_handle* h, const double* r, double* z);
/* GPU-free view of what the device runs, for host-logic tests: stage 0 / 1 = first / second triangular solve.  Row i of a
 * stage: t = x[i]; t -= values[k] * x[colidx[k]] over its entries in order (one FMA per step when `fused`, else product
 * and subtraction rounded separately -- whichever the reference's compiled loop does); x[i] = unit_diag ? t : t / diag[i].
 * level_rows[level_ptr[l] .. level_ptr[l+1]) are the rows of level l; launches holds (level_begin, level_end, widest
 * level) per kernel launch.  permscale: x[k] = pre_scale[k] * r[pre_gather[k]] before, z[k] = post_scale[k] *
 * x[post_gather[k]] after; present[0..3] says which of the four arrays exist (absent gather = identity, scale = 1). */
int b200s_factors_stage_sizes(const b200s_factors* f, int which, int64_t* nnz, int32_t* levels, int32_t* launches,
                              int32_t* unit_diag, int32_t* fused);
int b200s_factors_stage(const b200s_factors* f, int which, int32_t* rowptr, int32_t* colidx, double* values, double* diag,
                        int32_t* level_ptr, int32_t* level_rows, int32_t* launches);
int b200s_factors_permscale(const b200s_factors* f, int32_t* pre_gather, double* pre_scale, int32_t* post_gather,
                            double* post_scale, int32_t* present);

/* ---- introspection --------------------------------------------------------------------------------------------- */
int b200s_get_stats(b200s_handle* h, b200s_stats* out /* out->struct_size must be set */);
/* Copies the preconditioner's inverse diagonal (this rank's rows) to the host: DiagonalPreconditioner::m_invdiag. */
int b200s_get_invdiag_f64(b200s_handle* h, double* invdiag);
/* Per-iteration squared residual norms of the last solve (at most `cap`), for trajectory parity;
 * returns the number written, or a negative status.  Recorded by the fused CG / BiCGSTAB loops (Jacobi / identity);
 * 0 after a solve with an incomplete-factorization preconditioner. */
int64_t b200s_get_residual_history(b200s_handle* h, double* rr, int64_t cap);

/* Device-side timeline of the last solve (globaltimer, this rank): out[e] / out[12+e] = microseconds / count of the
 * interval ending at reduction kind e (1 spmv-only, 2 cg-init, 3 p.Ap, 4 cg-update, 5 bicg-init, 6 r0.v, 7 t.s/t.t,
 * 8 bicg-update, 9 restart, 10 end of the CG direction pass), each including the launch gap before it; out[24] = time inside cross-rank all-reduces,
 * out[25] = time CTA 0 waited for halo entries, out[26] = first-to-last reduction.  cap >= 27. */
int b200s_get_timeline(b200s_handle* h, double* out, int cap);

/* ---- GPU-free planning, exposed for host-logic tests (no CUDA call is made) ----------------------------------------
 * Builds the same partition / halo / tile plan analyze_pattern builds and reports it.  `local_colidx` (nnz entries,
 * optional) receives the remapped column indices: owned column c -> c - row_starts[rank]; ghost g -> rows + g.
 * `ghost_cols` (capacity ghost_cap, optional) receives the sorted global ids of the ghosts; `send_rows` (capacity
 * send_cap, optional) the local rows this rank sends, grouped by destination rank; `send_counts` / `recv_counts`
 * (world entries each, optional) the per-peer counts.  Returns the number of ghosts, or a negative status. */
int64_t b200s_plan_probe(const b200s_config* cfg, int64_t rows, int64_t cols, int64_t nnz, const int32_t* rowptr,
                         const int32_t* colidx, const int64_t* row_starts, int32_t* local_colidx,
                         int64_t* ghost_cols, int64_t ghost_cap, int32_t* send_rows, int64_t send_cap,
                         int64_t* send_counts, int64_t* recv_counts, b200s_stats* tile_stats);

/* GPU-free view of the canonical device matrix analyze_pattern would build on ONE rank from an uncompressed and/or
 * one-triangle input: compressed CSR, symmetric expansion for LOWER / UPPER (what selfadjointView implies,
 * SparseSelfAdjointView.h:279-337).  out_rowptr has rows+1 entries; out_colidx / out_src (optional) receive up to `cap`
 * entries, out_src[k] being the index into the caller's value array that entry k takes its value from.  Returns the
 * number of entries of the canonical matrix, or a negative status. */
int64_t b200s_plan_probe_csr(int64_t rows, int64_t nnz, const int32_t* rowptr, const int32_t* colidx,
                             const int32_t* inner_nnz, int uplo, int32_t* out_rowptr, int32_t* out_colidx,
                             int32_t* out_src, int64_t cap);

/* GPU-free view of THIS RANK's rows of the device matrix analyze_pattern + factorize_f64 build from a row-partitioned
 * input, in particular from one stored triangle (uplo LOWER / UPPER with world > 1: the mirror image of an entry stored
 * by another rank is fetched through the config's allgather, pattern and values alike).  A collective when world > 1:
 * every rank calls it.  out_rowptr has rows+1 entries; out_cols (GLOBAL column ids) and out_values (optional, needs
 * `values`) receive up to `cap` entries in the order the SpMV kernel sums them.  Returns the number of local entries
 * or a negative status. */
int64_t b200s_plan_probe_selfadjoint(const b200s_config* cfg, int64_t rows, int64_t cols, int64_t nnz,
                                     const int32_t* rowptr, const int32_t* colidx, const int32_t* inner_nnz, int uplo,
                                     const int64_t* row_starts, const double* values, int32_t* out_rowptr,
                                     int64_t* out_cols, double* out_values, int64_t cap);

/* Number of leading slots of the caller's value array that factorize will read for this pattern (one past the last
 * slot any row references), or a negative status.  GPU-free. */
int64_t b200s_plan_probe_span(int64_t rows, int64_t nnz, const int32_t* rowptr, const int32_t* colidx,
                              const int32_t* inner_nnz, int uplo);

#ifdef __cplusplus
}
#endif
#endif /* B200SPARSE_H */

"""Host-side mirror of the reference's solver interface for the B200 path.

``ConjugateGradient`` and ``BiCGSTAB`` keep the names, argument meaning and error behaviour of
/root/reference/Eigen/src/IterativeLinearSolvers/IterativeSolverBase.h:142-440 (compute / analyzePattern / factorize /
solve / solveWithGuess / setTolerance / tolerance / setMaxIterations / maxIterations / iterations / error / info) and
of ConjugateGradient.h:157-225 / BiCGSTAB.h:157-208, and forward to the C ABI of include/b200sparse.h.  The C++
counterpart for Eigen users is include/b200/IterativeSolvers.h.  There is no CPU path: everything numeric happens in
libb200sparse.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import B200Error, Config, Stats
from .workloads import CsrMatrix

# Eigen's enums (Core/util/Constants.h)
Lower, Upper = 1, 2
Success, NumericalIssue, NoConvergence, InvalidInput = 0, 1, 2, 3
IdentityPreconditioner, DiagonalPreconditioner = 0, 1

SPMV_AUTO, SPMV_STAGED, SPMV_DIRECT = 0, 1, 2
LOOP_AUTO, LOOP_WHILE_GRAPH, LOOP_CHUNKED_GRAPH, LOOP_STREAM, LOOP_PERSISTENT = 0, 1, 2, 3, 4


def device_count() -> int:
    return _lib.lib().b200s_device_count()


def _ptr(a) -> C.c_void_p:
    """Pointer of a numpy array, a torch tensor (data_ptr) or a raw integer address."""
    if a is None:
        return C.c_void_p(None)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(int(a))


def _index32(a, what: str) -> np.ndarray:
    """The C ABI reads int32 indices through raw pointers: hand it nothing else (Eigen's default StorageIndex)."""
    a = np.asarray(a)
    if a.dtype == np.int32 and a.flags.c_contiguous:
        return a
    if a.size and (a.max() > np.iinfo(np.int32).max or a.min() < np.iinfo(np.int32).min):
        raise ValueError(f"{what} does not fit int32 (StorageIndex = int)")
    return np.ascontiguousarray(a, dtype=np.int32)


def _as_csr(A) -> CsrMatrix:
    if isinstance(A, CsrMatrix):
        rp, ci = _index32(A.rowptr, "rowptr"), _index32(A.colidx, "colidx")
        vals = np.ascontiguousarray(A.vals)
        if rp is A.rowptr and ci is A.colidx and vals is A.vals:
            return A
        return CsrMatrix(A.rows, A.cols, rp, ci, vals, A.row0, A.name)
    if hasattr(A, "indptr"):  # scipy.sparse
        A = A.tocsr()
        return CsrMatrix(A.shape[0], A.shape[1], _index32(A.indptr, "rowptr"), _index32(A.indices, "colidx"),
                         np.ascontiguousarray(A.data), 0)
    raise TypeError("expected a CsrMatrix or a scipy.sparse matrix")


class Communicator:
    """Row-block partition context of one process per GPU.  ``allgather(bytes) -> list[bytes]`` is only used while
    the plan is built (setup); the iteration itself talks over NVLink peer memory."""

    def __init__(self, rank: int, world: int, allgather, row_starts):
        self.rank, self.world = int(rank), int(world)
        self._allgather = allgather
        self.row_starts = np.ascontiguousarray(row_starts, dtype=np.int64)
        assert self.row_starts.shape[0] == self.world + 1
        self.errors = []

        def cb(_ctx, send, recv, nbytes):
            try:
                mine = C.string_at(send, nbytes)
                parts = self._allgather(mine)
                assert len(parts) == self.world and all(len(p) == nbytes for p in parts)
                C.memmove(recv, b"".join(parts), nbytes * self.world)
                return 0
            except Exception as e:  # never let an exception cross the C boundary
                self.errors.append(e)
                return 1

        self.callback = _lib.ALLGATHER_FN(cb)

    @staticmethod
    def from_torch(row_starts, group=None):
        """Bootstrap over torch.distributed (a gloo group; with an NCCL default group pass a gloo subgroup)."""
        import torch
        import torch.distributed as dist

        def allgather(b: bytes):
            t = torch.frombuffer(bytearray(b), dtype=torch.uint8)
            out = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
            dist.all_gather(out, t, group=group)
            return [bytes(o.numpy().tobytes()) for o in out]

        return Communicator(dist.get_rank(group), dist.get_world_size(group), allgather, row_starts)


def partition_rows(n: int, world: int, align: int = 1, rowptr=None, vector_bytes_per_row: int = 104) -> np.ndarray:
    """Contiguous row blocks; ``align`` keeps block edges on grid-plane boundaries.

    Without ``rowptr`` the blocks hold equal numbers of rows (right for stencils).  With the GLOBAL ``rowptr`` they
    are balanced by the bytes one solver iteration streams per row -- 12 per stored entry plus
    ``vector_bytes_per_row`` (CG: 13 vector passes of 8 bytes) -- which is what keeps ranks in step on matrices with
    uneven row lengths (power-law)."""
    if rowptr is None:
        units = n // align
        starts = [(units * r // world) * align for r in range(world)] + [n]
        return np.asarray(starts, dtype=np.int64)
    rp = np.asarray(rowptr, dtype=np.int64)
    assert rp.shape[0] == n + 1
    cost = 12 * (rp - rp[0]) + vector_bytes_per_row * np.arange(n + 1, dtype=np.int64)
    targets = cost[-1] * np.arange(1, world, dtype=np.float64) / world
    cuts = np.searchsorted(cost, targets, side="left")
    cuts = (np.round(cuts / align).astype(np.int64) * align).clip(0, n)
    starts = np.concatenate([[0], np.maximum.accumulate(cuts), [n]])
    return starts.astype(np.int64)


class _Handle:
    """Owns one b200s_handle."""

    def __init__(self, comm: Optional[Communicator] = None, device: int = -1, spmv_impl: int = 0, loop_mode: int = 0,
                 chunk_iters: int = 0, tile_nnz: int = 0, tile_rows: int = 0):
        self.L = _lib.lib()
        self.comm = comm
        cfg = Config()
        cfg.struct_size = C.sizeof(Config)
        cfg.device = device
        cfg.rank = comm.rank if comm else 0
        cfg.world = comm.world if comm else 1
        cfg.spmv_impl, cfg.loop_mode, cfg.chunk_iters = spmv_impl, loop_mode, chunk_iters
        cfg.tile_nnz, cfg.tile_rows = tile_nnz, tile_rows
        if comm:
            cfg.allgather = comm.callback
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.L.b200s_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise B200Error(rc, self.L.b200s_last_error(None).decode())

    def check(self, rc):
        if rc != 0:
            msg = self.L.b200s_last_error(self.h).decode()
            if self.comm and self.comm.errors:
                msg += f" (allgather callback: {self.comm.errors[-1]!r})"
            raise B200Error(rc, msg)

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.b200s_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self) -> dict:
        st = Stats()
        st.struct_size = C.sizeof(Stats)
        self.check(self.L.b200s_get_stats(self.h, C.byref(st)))
        return st.as_dict()


class SparseOperator:
    """``y = A * x`` on the GPU: the product of SparseDenseProduct.h:26-72 for a row-major matrix."""

    def __init__(self, A=None, uplo: int = Lower | Upper, comm: Optional[Communicator] = None, **cfg):
        self._hd = _Handle(comm, **cfg)
        self._comm = comm
        self._uplo = uplo
        self._dtype = None
        self._rows = self._cols = 0
        if A is not None:
            self.compute(A)

    def analyzePattern(self, A, inner_nnz=None):
        A = _as_csr(A)
        self._rows, self._cols = A.rows, A.cols
        rs = self._comm.row_starts if self._comm else None
        inz = None if inner_nnz is None else np.ascontiguousarray(inner_nnz, np.int32)
        self._keep = (A.rowptr, A.colidx, inz, rs)
        nnz = int(A.colidx.shape[0])
        self._hd.check(self._hd.L.b200s_analyze_pattern(self._hd.h, A.rows, A.cols, nnz, _ptr(A.rowptr), _ptr(A.colidx),
                                                        _ptr(inz), self._uplo, _ptr(rs)))
        return self

    def factorize(self, A, precond: int = DiagonalPreconditioner):
        A = _as_csr(A)
        vals = np.ascontiguousarray(A.vals)
        if vals.dtype == np.float32:
            self._hd.check(self._hd.L.b200s_factorize_f32(self._hd.h, _ptr(vals), precond))
        elif vals.dtype == np.float64:
            self._hd.check(self._hd.L.b200s_factorize_f64(self._hd.h, _ptr(vals), precond))
        else:
            raise TypeError("values must be float32 or float64")
        self._dtype = vals.dtype
        return self

    def compute(self, A, precond: int = DiagonalPreconditioner, inner_nnz=None):
        return self.analyzePattern(A, inner_nnz).factorize(A, precond)

    def rows(self):
        return self._rows

    def cols(self):
        return self._cols

    def multiply(self, x: np.ndarray) -> np.ndarray:
        """Host vectors in, host vector out (H2D + kernel + D2H)."""
        x = np.ascontiguousarray(x, dtype=self._dtype)
        nx = self._cols if not self._comm or self._comm.world == 1 else self._rows
        if x.shape != (nx,):
            raise ValueError(f"x must have shape ({nx},)")
        y = np.empty(self._rows, dtype=self._dtype)
        fn = self._hd.L.b200s_spmv_f32 if self._dtype == np.float32 else self._hd.L.b200s_spmv_f64
        self._hd.check(fn(self._hd.h, _ptr(x), _ptr(y)))
        return y

    __matmul__ = multiply

    def multiply_device(self, x_dev, y_dev, reps: int = 1) -> float:
        """Device pointers (torch tensors or addresses); returns the average kernel time in ms."""
        ms = C.c_float(0)
        fn = self._hd.L.b200s_spmv_device_f32 if self._dtype == np.float32 else self._hd.L.b200s_spmv_device_f64
        self._hd.check(fn(self._hd.h, _ptr(x_dev), _ptr(y_dev), reps, C.byref(ms)))
        return ms.value

    def invdiag(self) -> np.ndarray:
        """DiagonalPreconditioner::m_invdiag of a double factorization."""
        d = np.empty(self._rows, dtype=np.float64)
        self._hd.check(self._hd.L.b200s_get_invdiag_f64(self._hd.h, _ptr(d)))
        return d

    def stats(self) -> dict:
        return self._hd.stats()

    def close(self):
        self._hd.close()


class _IterativeSolverBase(SparseOperator):
    """IterativeSolverBase.h:142-440."""

    _bicg = False

    def __init__(self, A=None, uplo: int = Lower | Upper, preconditioner=DiagonalPreconditioner,
                 comm: Optional[Communicator] = None, **cfg):
        # preconditioner: IdentityPreconditioner / DiagonalPreconditioner, or an IncompleteLUT / IncompleteCholesky
        # object (preconditioners.py) -- the template argument of the reference's solver classes
        self._pre_obj = None
        if not isinstance(preconditioner, (int, np.integer)):
            self._pre_obj = preconditioner
            preconditioner = DiagonalPreconditioner  # what factorize() builds underneath; replaced right after
        self._precond = int(preconditioner)
        self._tolerance = -1.0                               # :413 -> NumTraits<Scalar>::epsilon(), resolved per dtype
        self._max_iterations = -1                            # :281-284 -> 2*cols
        self._iterations = 0
        self._error = 0.0
        self._info = Success
        self._is_initialized = False
        super().__init__(None, uplo, comm, **cfg)
        if A is not None:
            self.compute(A)

    # ---- setup (IterativeSolverBase.h:196-247) ----
    def analyzePattern(self, A, inner_nnz=None):
        super().analyzePattern(A, inner_nnz)
        self._analysis_ok = True
        self._is_initialized = True
        self._info = Success
        return self

    def factorize(self, A, precond=None):
        if not getattr(self, "_analysis_ok", False):
            raise AssertionError("You must first call analyzePattern()")  # :218
        super().factorize(A, self._precond if precond is None else precond)
        self._factorization_ok = True
        self._info = Success
        if self._pre_obj is not None:
            # IterativeSolverBase::factorize: m_preconditioner.factorize(matrix()); m_info = m_preconditioner.info()
            # (IterativeSolverBase.h:216-224).  The factorization is host setup; the solves apply it on the GPU.
            self._pre_obj.compute(A)
            self._info = self._pre_obj.info()
            if self._info == Success:
                self._hd.check(self._hd.L.b200s_set_preconditioner(self._hd.h, self._pre_obj.handle()))
        return self

    def preconditioner(self):
        """The preconditioner object (IterativeSolverBase.h:249-253), or the enum value for Identity / Diagonal."""
        return self._pre_obj if self._pre_obj is not None else self._precond

    def precondition(self, r: np.ndarray) -> np.ndarray:
        """z = preconditioner().solve(r) on the GPU (host vectors; for tests and diagnostics)."""
        r = np.ascontiguousarray(r, dtype=np.float64)
        if r.shape != (self._rows,):
            raise ValueError(f"r must have shape ({self._rows},)")
        z = np.empty_like(r)
        self._hd.check(self._hd.L.b200s_precond_apply_f64(self._hd.h, _ptr(r), _ptr(z)))
        return z

    def compute(self, A, precond=None, inner_nnz=None):
        self.analyzePattern(A, inner_nnz)
        return self.factorize(A, precond)

    # ---- parameters (:258-293) ----
    def tolerance(self):
        if self._tolerance < 0:
            return float(np.finfo(self._dtype if self._dtype is not None else np.float64).eps)
        return self._tolerance

    def setTolerance(self, tol):
        self._tolerance = float(tol)
        return self

    def maxIterations(self):
        return 2 * self._cols if self._max_iterations < 0 else self._max_iterations

    def setMaxIterations(self, n):
        self._max_iterations = int(n)
        return self

    # ---- results (:296-330) ----
    def iterations(self):
        if not self._is_initialized:
            raise AssertionError("ConjugateGradient is not initialized.")
        return self._iterations

    def error(self):
        if not self._is_initialized:
            raise AssertionError("ConjugateGradient is not initialized.")
        return self._error

    def info(self):
        if not self._is_initialized:
            raise AssertionError("IterativeSolverBase is not initialized.")
        return self._info

    # ---- solves (:316-323, :333-404) ----
    def _solve_vector(self, b: np.ndarray, x: np.ndarray, use_guess: bool):
        it, err, info = C.c_int64(0), C.c_double(0), C.c_int(0)
        fn = getattr(self._hd.L, f"b200s_{'bicgstab' if self._bicg else 'cg'}_solve_{self._sfx()}")
        self._hd.check(fn(self._hd.h, _ptr(b), _ptr(x), int(use_guess), self.tolerance(), self.maxIterations(),
                          C.byref(it), C.byref(err), C.byref(info)))
        return it.value, err.value, info.value

    def _sfx(self):
        return "f32" if self._dtype == np.float32 else "f64"

    def _solve(self, b, x0):
        if not self._is_initialized:
            raise AssertionError("solver is not initialized.")  # :337
        dt = self._dtype if self._dtype is not None else np.float64
        b = np.asarray(b, dtype=dt)
        if b.shape[0] != self._rows:
            raise AssertionError("solve(): invalid number of rows of the right hand side matrix b")
        if x0 is not None and np.shape(x0) != b.shape:
            raise AssertionError("solveWithGuess(): the guess must have the shape of the right hand side")  # :319
        if b.ndim == 1:
            x = np.zeros(self._rows, dt) if x0 is None else np.array(x0, dtype=dt, copy=True)
            self._iterations, self._error, self._info = self._solve_vector(np.ascontiguousarray(b), x, x0 is not None)
            return x
        # multi-column right-hand side: info = worst, iterations / error = last column (:375-388)
        X = np.zeros(b.shape, dt, order="F") if x0 is None else np.array(x0, dtype=dt, order="F", copy=True)
        global_info = Success
        if not self._bicg and dt == np.float64 and b.shape[1] > 0:
            # CG: the columns share one stream of the matrix per iteration (b200s_cg_solve_multi_f64); per-column
            # results are bit-identical to the sequential loop of the reference
            Bf = np.asfortranarray(b)
            nc = b.shape[1]
            its, errs, infos = np.zeros(nc, np.int64), np.zeros(nc, np.float64), np.zeros(nc, np.int32)
            self._hd.check(self._hd.L.b200s_cg_solve_multi_f64(
                self._hd.h, nc, _ptr(Bf), max(1, Bf.shape[0]), _ptr(X), max(1, X.shape[0]), int(x0 is not None),
                self.tolerance(), self.maxIterations(), _ptr(its), _ptr(errs), _ptr(infos)))
            self.column_iterations, self.column_errors, self.column_infos = its, errs, infos
            for info in infos:
                if info == NumericalIssue:
                    global_info = NumericalIssue
                elif info == NoConvergence:
                    global_info = NoConvergence
            self._iterations, self._error, self._info = int(its[-1]), float(errs[-1]), global_info
            return X
        for k in range(b.shape[1]):
            xk = np.ascontiguousarray(X[:, k])
            self._iterations, self._error, info = self._solve_vector(np.ascontiguousarray(b[:, k]), xk, x0 is not None)
            X[:, k] = xk
            # IterativeSolverBase.h:355-358 / :383-386, literally: a later NoConvergence overwrites NumericalIssue
            if info == NumericalIssue:
                global_info = NumericalIssue
            elif info == NoConvergence:
                global_info = NoConvergence
        self._info = global_info
        return X

    def solve(self, b):
        return self._solve(b, None)

    def solveWithGuess(self, b, x0):
        return self._solve(b, x0)

    def solve_device(self, b_dev, x_dev, use_guess: bool = False):
        """Inputs resident in HBM (torch tensors / device addresses of this rank's rows)."""
        it, err, info = C.c_int64(0), C.c_double(0), C.c_int(0)
        fn = getattr(self._hd.L, f"b200s_{'bicgstab' if self._bicg else 'cg'}_solve_device_{self._sfx()}")
        self._hd.check(fn(self._hd.h, _ptr(b_dev), _ptr(x_dev), int(use_guess), self.tolerance(), self.maxIterations(),
                          C.byref(it), C.byref(err), C.byref(info)))
        self._iterations, self._error, self._info = it.value, err.value, info.value
        return x_dev

    def solve_device_multi(self, B_dev, X_dev, ncols: int, ldb: int = 0, ldx: int = 0, use_guess: bool = False):
        """CG with `ncols` right-hand sides resident in HBM, column-major (leading dimensions default to rows)."""
        its, errs, infos = np.zeros(ncols, np.int64), np.zeros(ncols, np.float64), np.zeros(ncols, np.int32)
        self._hd.check(self._hd.L.b200s_cg_solve_multi_device_f64(
            self._hd.h, ncols, _ptr(B_dev), ldb or self._rows, _ptr(X_dev), ldx or self._rows, int(use_guess),
            self.tolerance(), self.maxIterations(), _ptr(its), _ptr(errs), _ptr(infos)))
        self.column_iterations, self.column_errors, self.column_infos = its, errs, infos
        self._iterations, self._error = int(its[-1]), float(errs[-1])
        self._info = NoConvergence if (infos == NoConvergence).any() else Success
        return X_dev

    def multi_rhs_batch(self) -> int:
        """Columns that share one stream of the matrix per iteration (0: this handle solves column by column)."""
        return int(self._hd.L.b200s_multi_rhs_batch(self._hd.h))

    def timeline(self) -> dict:
        """Device-side timeline of the last solve on this rank (see b200s_get_timeline)."""
        out = np.zeros(27)
        self._hd.check(self._hd.L.b200s_get_timeline(self._hd.h, _ptr(out), 27))
        names = {1: "spmv_only", 2: "cg_init", 3: "spmv_pAp", 4: "cg_update", 5: "bicg_init", 6: "spmv_r0v",
                 7: "spmv_ts_tt", 8: "bicg_update", 9: "bicg_restart", 10: "cg_direction"}
        d = {f"{names[e]}_us_avg": out[e] / out[12 + e] for e in names if out[12 + e] > 0}
        d.update(allreduce_us_total=out[24], halo_wait_us_total=out[25], span_us=out[26],
                 reductions=int(out[12:24].sum()))
        return d

    def residual_history(self, cap: int = 1 << 16) -> np.ndarray:
        rr = np.empty(cap, dtype=np.float64)
        n = self._hd.L.b200s_get_residual_history(self._hd.h, _ptr(rr), cap)
        if n < 0:
            self._hd.check(int(n))
        return rr[:n].copy()


class ConjugateGradient(_IterativeSolverBase):
    """ConjugateGradient<SparseMatrix<Scalar,RowMajor>, UpLo, Preconditioner> (ConjugateGradient.h:157-225); Scalar =
    double or float follows the dtype of the matrix values given to compute()/factorize()."""
    _bicg = False


class BiCGSTAB(_IterativeSolverBase):
    """BiCGSTAB<SparseMatrix<Scalar,RowMajor>, Preconditioner> (BiCGSTAB.h:157-208); the matrix is used as stored."""
    _bicg = True

    def __init__(self, A=None, preconditioner=DiagonalPreconditioner, comm: Optional[Communicator] = None, **cfg):
        super().__init__(A, Lower | Upper, preconditioner, comm, **cfg)


class MINRES(_IterativeSolverBase):
    """MINRES<SparseMatrix<double>, UpLo, Preconditioner> (unsupported/Eigen/src/IterativeSolvers/MINRES.h:29-262) for
    self-adjoint operators; the default preconditioner is the identity, as in the reference.  One GPU, double."""

    def __init__(self, A=None, uplo: int = Lower, preconditioner=IdentityPreconditioner, **cfg):
        super().__init__(A, uplo, preconditioner, None, **cfg)

    def _solve_vector(self, b, x, use_guess):
        it, err, info = C.c_int64(0), C.c_double(0), C.c_int(0)
        self._hd.check(self._hd.L.b200s_minres_solve_f64(self._hd.h, _ptr(b), _ptr(x), int(use_guess), self.tolerance(),
                                                         self.maxIterations(), C.byref(it), C.byref(err), C.byref(info)))
        return it.value, err.value, info.value

    _bicg = True  # multi-column right-hand sides: the reference's per-column loop


class GMRES(_IterativeSolverBase):
    """GMRES<SparseMatrix<double>, Preconditioner> (unsupported/Eigen/src/IterativeSolvers/GMRES.h:55-325): restarted,
    Householder Arnoldi; set_restart / get_restart as in the reference (default 30).  One GPU, double."""
    _bicg = True

    def __init__(self, A=None, preconditioner=DiagonalPreconditioner, **cfg):
        self._restart = 30
        super().__init__(A, Lower | Upper, preconditioner, None, **cfg)

    def get_restart(self):
        return self._restart

    def set_restart(self, restart):
        self._restart = int(restart)

    def _solve_vector(self, b, x, use_guess):
        it, err, info = C.c_int64(0), C.c_double(0), C.c_int(0)
        self._hd.check(self._hd.L.b200s_gmres_solve_f64(self._hd.h, _ptr(b), _ptr(x), int(use_guess), self.tolerance(),
                                                        self.maxIterations(), self._restart, C.byref(it), C.byref(err),
                                                        C.byref(info)))
        return it.value, err.value, info.value


class LeastSquaresConjugateGradient(_IterativeSolverBase):
    """LeastSquaresConjugateGradient<SparseMatrix<double,RowMajor>, Preconditioner>
    (LeastSquareConjugateGradient.h:26-208): min |A x - b| for a rows x cols matrix.  The transposed matrix needed for
    A^T r is built once per compute() on the host (scipy) and lives in a second device handle.  preconditioner:
    DiagonalPreconditioner selects LeastSquareDiagonalPreconditioner (the reference's default), Identity the identity.
    One GPU, double."""
    _bicg = True

    def __init__(self, A=None, preconditioner: int = DiagonalPreconditioner, **cfg):
        self._t = None
        super().__init__(A, Lower | Upper, preconditioner, None, **cfg)

    def analyzePattern(self, A, inner_nnz=None):
        A = _as_csr(A)
        super().analyzePattern(A, inner_nnz)
        At = A.to_scipy().T.tocsr()
        At.sort_indices()
        self._At = CsrMatrix(A.cols, A.rows, _index32(At.indptr, "rowptr"), _index32(At.indices, "colidx"),
                             np.ascontiguousarray(At.data, dtype=np.float64), 0)
        if self._t is None:
            self._t = SparseOperator()
        self._t.analyzePattern(self._At)
        return self

    def factorize(self, A, precond=None):
        A = _as_csr(A)
        super().factorize(A, IdentityPreconditioner)  # the LS preconditioner is built from A^T inside the solve
        At = A.to_scipy().T.tocsr()
        At.sort_indices()
        self._At = CsrMatrix(A.cols, A.rows, self._At.rowptr, self._At.colidx, np.ascontiguousarray(At.data, np.float64), 0)
        self._t.factorize(self._At, IdentityPreconditioner)
        return self

    def maxIterations(self):
        return 2 * self._cols if self._max_iterations < 0 else self._max_iterations

    def _solve(self, b, x0):
        if not self._is_initialized:
            raise AssertionError("solver is not initialized.")
        b = np.ascontiguousarray(b, dtype=np.float64)
        if b.ndim != 1 or b.shape[0] != self._rows:
            raise AssertionError("solve(): invalid number of rows of the right hand side matrix b")
        x = np.zeros(self._cols) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
        if x.shape != (self._cols,):
            raise AssertionError("solveWithGuess(): the guess must have cols() entries")
        it, err, info = C.c_int64(0), C.c_double(0), C.c_int(0)
        self._hd.check(self._hd.L.b200s_lscg_solve_f64(self._hd.h, self._t._hd.h, _ptr(b), _ptr(x), int(x0 is not None),
                                                       self.tolerance(), self.maxIterations(), self._precond, 0,
                                                       C.byref(it), C.byref(err), C.byref(info)))
        self._iterations, self._error, self._info = it.value, err.value, info.value
        return x

    def close(self):
        if self._t is not None:
            self._t.close()
        super().close()

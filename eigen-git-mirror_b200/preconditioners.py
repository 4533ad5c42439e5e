"""IncompleteLUT / IncompleteCholesky for the B200 solvers.

Mirrors the reference's classes (Eigen/src/IterativeLinearSolvers/IncompleteLUT.h:98-190, IncompleteCholesky.h:40-170):
``compute`` factorizes on the host (sequential setup work, the same factors as the reference entry for entry for the
same permutation), and a solver that is given the object -- ``ConjugateGradient(A, preconditioner=IncompleteCholesky())``
-- applies it on the GPU in every iteration (level-scheduled triangular solves, csrc/kernels_tri.cuh).  The
fill-reducing permutation is an input (``perm``, in the reference's own convention m_P / m_perm ``.indices()``); None is
the natural ordering.  No numerical work happens in Python."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from ._lib import B200Error

Lower, Upper = 1, 2
FACTORS_ILUT, FACTORS_ICHOL = 1, 2


def _ptr(a):
    return C.c_void_p(None) if a is None else C.c_void_p(a.ctypes.data)


@dataclass
class Stage:
    """One triangular solve as the device runs it (GPU-free view, for host-logic tests)."""
    rowptr: np.ndarray
    colidx: np.ndarray
    vals: np.ndarray
    diag: Optional[np.ndarray]   # None = unit diagonal
    level_ptr: np.ndarray
    level_rows: np.ndarray
    launches: np.ndarray         # (n_launches, 3): level_begin, level_end, widest level
    fused: bool = False          # each step one FMA (else product and subtraction rounded separately)


def multicolor_ordering(A):
    """Permutation with few, wide dependency levels for the GPU's triangular solves: greedy multi-colouring of the
    pattern of A + A^T, rows sorted by colour (b200s_ordering_multicolor; host, GPU-free).  Returns (perm, colours);
    perm is what IncompleteLUT / IncompleteCholesky take as ``perm``."""
    from .solvers import _as_csr
    A = _as_csr(A)
    perm = np.zeros(A.rows, np.int32)
    L = _lib.lib()
    nc = L.b200s_ordering_multicolor(A.rows, _ptr(A.rowptr), _ptr(A.colidx), _ptr(perm))
    if nc < 0:
        raise B200Error(nc, L.b200s_last_error(None).decode())
    return perm, int(nc)


class _Factors:
    """Owns one b200s_factors object."""

    kind = 0

    def __init__(self):
        self.L = _lib.lib()
        self._f = C.c_void_p()

    def _take(self, rc):
        if rc != 0:
            raise B200Error(rc, self.L.b200s_last_error(None).decode())
        return self

    def _reset(self):
        if self._f.value:
            self.L.b200s_factors_destroy(self._f)
        self._f = C.c_void_p()

    def __del__(self):
        try:
            self._reset()
        except Exception:
            pass

    # ---- the reference's accessors ----
    def info(self) -> int:
        if not self._f.value:
            raise AssertionError(f"{type(self).__name__} is not initialized.")
        return int(self.L.b200s_factors_info(self._f))

    def rows(self) -> int:
        return int(self.L.b200s_factors_size(self._f)) if self._f.value else 0

    cols = rows

    def handle(self):
        if not self._f.value:
            raise AssertionError(f"{type(self).__name__} is not initialized.")
        return self._f

    def arrays(self):
        """(outer, inner, values, scale, perm): m_lu resp. m_L / m_scale / m_perm as the reference stores them."""
        n, nz, ps = self.rows(), int(self.L.b200s_factors_nnz(self._f)), int(self.L.b200s_factors_perm_size(self._f))
        outer, inner, vals = np.zeros(n + 1, np.int32), np.zeros(nz, np.int32), np.zeros(nz, np.float64)
        scale = np.zeros(n if self.kind == FACTORS_ICHOL else 0, np.float64)
        perm = np.zeros(ps, np.int32)
        self.L.b200s_factors_get(self._f, _ptr(outer), _ptr(inner) if nz else None, _ptr(vals) if nz else None,
                                 _ptr(scale) if scale.size else None, _ptr(perm) if ps else None)
        return outer, inner, vals, scale, perm

    def stage(self, which: int) -> Stage:
        n = self.rows()
        nnz, lev, lau, unit, fused = C.c_int64(0), C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._take(self.L.b200s_factors_stage_sizes(self._f, which, C.byref(nnz), C.byref(lev), C.byref(lau),
                                                    C.byref(unit), C.byref(fused)))
        rp, ci, va = np.zeros(n + 1, np.int32), np.zeros(max(1, nnz.value), np.int32), np.zeros(max(1, nnz.value))
        dg = None if unit.value else np.zeros(max(1, n))
        lp, lr = np.zeros(lev.value + 1, np.int32), np.zeros(max(1, n), np.int32)
        la = np.zeros((max(1, lau.value), 3), np.int32)
        self._take(self.L.b200s_factors_stage(self._f, which, _ptr(rp), _ptr(ci), _ptr(va), _ptr(dg), _ptr(lp), _ptr(lr),
                                              _ptr(la)))
        return Stage(rp, ci[:nnz.value], va[:nnz.value], None if dg is None else dg[:n], lp, lr[:n], la[:lau.value],
                     bool(fused.value))

    def permscale(self):
        """(pre_gather, pre_scale, post_gather, post_scale), None where absent."""
        n = self.rows()
        present = np.zeros(4, np.int32)
        self.L.b200s_factors_permscale(self._f, None, None, None, None, _ptr(present))
        bufs = [np.zeros(max(1, n), np.int32 if i % 2 == 0 else np.float64) if present[i] else None for i in range(4)]
        self.L.b200s_factors_permscale(self._f, *[_ptr(b) for b in bufs], None)
        return tuple(None if b is None else b[:n] for b in bufs)


class IncompleteLUT(_Factors):
    """IncompleteLUT<double> (IncompleteLUT.h:98-190): dual-threshold ILU, ``droptol`` and ``fillfactor`` as there."""

    kind = FACTORS_ILUT

    def __init__(self, A=None, droptol: float = -1.0, fillfactor: int = 0, perm=None):
        super().__init__()
        self._droptol, self._fillfactor = float(droptol), int(fillfactor)
        self._perm = None if perm is None else np.ascontiguousarray(perm, np.int32)
        if A is not None:
            self.compute(A)

    def setDroptol(self, droptol):
        self._droptol = float(droptol)

    def setFillfactor(self, fillfactor):
        self._fillfactor = int(fillfactor)

    def setPermutation(self, perm):
        """m_P.indices(): the reference computes it with AMD in analyzePattern (:221-236); here it is an input."""
        self._perm = None if perm is None else np.ascontiguousarray(perm, np.int32)

    def compute(self, A):
        from .solvers import _as_csr
        A = _as_csr(A)
        if A.rows != A.cols:
            raise AssertionError("The factorization should be done on a square matrix")  # IncompleteLUT.h:250
        self._reset()
        vals = np.ascontiguousarray(A.vals, np.float64)
        return self._take(self.L.b200s_ilut_f64(A.rows, _ptr(A.rowptr), _ptr(A.colidx), _ptr(vals), self._droptol,
                                                self._fillfactor, _ptr(self._perm), C.byref(self._f)))

    @classmethod
    def from_factors(cls, lu_rowptr, lu_colidx, lu_vals, perm=None):
        """Wrap a factor computed elsewhere (an Eigen::IncompleteLUT's m_lu / m_P)."""
        self = cls()
        rp, ci = np.ascontiguousarray(lu_rowptr, np.int32), np.ascontiguousarray(lu_colidx, np.int32)
        va = np.ascontiguousarray(lu_vals, np.float64)
        pm = None if perm is None else np.ascontiguousarray(perm, np.int32)
        return self._take(self.L.b200s_factors_from_ilut_f64(rp.shape[0] - 1, _ptr(rp), _ptr(ci), _ptr(va), _ptr(pm),
                                                             C.byref(self._f)))


class IncompleteCholesky(_Factors):
    """IncompleteCholesky<double, UpLo, Ordering> (IncompleteCholesky.h:40-170); ``uplo`` = the triangle that is read."""

    kind = FACTORS_ICHOL

    def __init__(self, A=None, uplo: int = Lower, perm=None, initial_shift: float = -1.0):
        super().__init__()
        self._uplo, self._shift = int(uplo), float(initial_shift)
        self._perm = None if perm is None else np.ascontiguousarray(perm, np.int32)
        if A is not None:
            self.compute(A)

    def setInitialShift(self, shift):
        self._shift = float(shift)

    def setPermutation(self, perm):
        """m_perm.indices() (the inverse of what the ordering functor returns, :95-105); None = NaturalOrdering."""
        self._perm = None if perm is None else np.ascontiguousarray(perm, np.int32)

    def compute(self, A):
        from .solvers import _as_csr
        A = _as_csr(A)
        self._reset()
        vals = np.ascontiguousarray(A.vals, np.float64)
        return self._take(self.L.b200s_ichol_f64(A.rows, _ptr(A.rowptr), _ptr(A.colidx), _ptr(vals), self._uplo,
                                                 self._shift, _ptr(self._perm), C.byref(self._f)))

    @classmethod
    def from_factors(cls, colptr, rowidx, lvals, scale=None, perm=None):
        self = cls()
        cp, ri = np.ascontiguousarray(colptr, np.int32), np.ascontiguousarray(rowidx, np.int32)
        va = np.ascontiguousarray(lvals, np.float64)
        sc = None if scale is None else np.ascontiguousarray(scale, np.float64)
        pm = None if perm is None or len(perm) == 0 else np.ascontiguousarray(perm, np.int32)
        return self._take(self.L.b200s_factors_from_ichol_f64(cp.shape[0] - 1, _ptr(cp), _ptr(ri), _ptr(va), _ptr(sc),
                                                              _ptr(pm), C.byref(self._f)))

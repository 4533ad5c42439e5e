"""MatrixMarket I/O with the semantics of the reference's SparseExtra module
(/root/reference/unsupported/Eigen/src/SparseExtra/MarketIO.h:109-282), the on-disk format of the reference's
real-matrix test flow (test/sparse_solver.h:147-179, bench/spbench/spbenchsolver.h:213-300).  SURVEY.md 8f rank 4.

* ``getMarketHeader``  -> (sym, iscomplex, isvector); sym uses Eigen's values Symmetric = 0x10, SelfAdjoint = 0x20
                          (Core/util/Constants.h), 0 for general.
* ``loadMarket``       -> CsrMatrix with the entries EXACTLY as stored: a symmetric file yields one triangle (the
                          caller picks UpLo, as spbenchsolver does), duplicates are summed (setFromTriplets),
                          out-of-range entries are skipped with a warning.
* ``saveMarket`` / ``loadMarketVector`` / ``saveMarketVector`` -- coordinate / array real, 17 significant digits.
* ``MatrixMarketIterator`` -- the folder walk of SparseExtra/MatrixMarketIterator.h:41-241 (name.mtx, name_b.mtx,
                          name_x.mtx, the "SPD" name convention), used by tools/solve_market.py.

Host-side only (numpy); nothing here touches the GPU.
"""
from __future__ import annotations

import sys

import numpy as np

from .workloads import CsrMatrix

Symmetric, SelfAdjoint = 0x10, 0x20


def getMarketHeader(filename: str):
    """MarketIO.h:109-131.  Returns (ok, sym, iscomplex, isvector)."""
    try:
        with open(filename) as f:
            line = f.readline()
    except OSError:
        return False, 0, False, False
    tok = (line.split() + [""] * 5)[:5]
    isvector = tok[2] == "array"
    iscomplex = tok[3] == "complex"
    sym = Symmetric if tok[4] == "symmetric" else SelfAdjoint if tok[4] == "Hermitian" else 0
    return True, sym, iscomplex, isvector


def loadMarket(filename: str, dtype=np.float64) -> CsrMatrix:
    """MarketIO.h:133-200: coordinate real matrix -> row-major CSR with sorted columns and summed duplicates."""
    rows = cols = nnz = -1
    ii, jj, vv = [], [], []
    with open(filename) as f:
        for line in f:
            if line.startswith("%") or not line.strip():
                continue
            parts = line.split()
            if rows < 0:
                m, n, z = int(parts[0]), int(parts[1]), int(parts[2])
                if m > 0 and n > 0 and z > 0:
                    rows, cols, nnz = m, n, z
                continue
            i, j = int(parts[0]) - 1, int(parts[1]) - 1
            v = float(parts[2]) if len(parts) > 2 else 1.0
            if 0 <= i < rows and 0 <= j < cols:
                ii.append(i); jj.append(j); vv.append(v)
            else:
                print(f"Invalid read: {i},{j}", file=sys.stderr)
    if rows < 0:
        raise ValueError(f"{filename}: no size line found")
    if len(ii) != nnz:
        print(f"{len(ii)}!={nnz}", file=sys.stderr)
    i = np.asarray(ii, np.int64); j = np.asarray(jj, np.int64); v = np.asarray(vv, dtype)
    order = np.lexsort((j, i))
    i, j, v = i[order], j[order], v[order]
    if i.size:
        first = np.ones(i.size, bool)
        first[1:] = (i[1:] != i[:-1]) | (j[1:] != j[:-1])
        grp = np.cumsum(first) - 1
        v = np.bincount(grp, weights=v).astype(dtype)
        i, j = i[first], j[first]
    counts = np.bincount(i, minlength=rows)
    rowptr = np.zeros(rows + 1, np.int32)
    np.cumsum(counts, out=rowptr[1:])
    return CsrMatrix(rows, cols, rowptr, j.astype(np.int32), v, 0, filename)


def saveMarket(A, filename: str, sym: int = 0) -> bool:
    """MarketIO.h:232-256: entries in storage order, 1-based, scientific with digits10+2 digits."""
    from .solvers import _as_csr
    A = _as_csr(A)
    kind = "symmetric" if sym == Symmetric else "Hermitian" if sym == SelfAdjoint else "general"
    try:
        with open(filename, "w") as out:
            out.write(f"%%MatrixMarket matrix coordinate  real {kind}\n")
            out.write(f"{A.rows} {A.cols} {A.nnz}\n")
            rowof = np.repeat(np.arange(A.rows), np.diff(A.rowptr))
            for r, c, v in zip(rowof, A.colidx, A.vals):
                out.write(f"{r + 1} {c + 1} {float(v):.16e}\n")
    except OSError:
        return False
    return True


def loadMarketVector(filename: str, dtype=np.float64) -> np.ndarray:
    """MarketIO.h:202-230: array format, first column only."""
    with open(filename) as f:
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        n, col = (int(t) for t in line.split()[:2])
        if n <= 0 or col <= 0:
            raise ValueError(f"{filename}: bad size line")
        vals = []
        for line in f:
            if len(vals) >= n:
                break
            if line.strip():
                vals.append(float(line.split()[0]))
    if len(vals) != n:
        raise ValueError(f"Unable to read all elements from file {filename}")
    return np.asarray(vals, dtype)


def saveMarketVector(vec, filename: str) -> bool:
    """MarketIO.h:258-280."""
    vec = np.asarray(vec, np.float64).ravel()
    try:
        with open(filename, "w") as out:
            out.write("%%MatrixMarket matrix array real general\n")
            out.write(f"{vec.size} 1\n")
            for v in vec:
                out.write(f"{float(v):.16e}\n")
    except OSError:
        return False
    return True


SPD, NonSymmetric = 0x100, 0x0


class MatrixMarketIterator:
    """MatrixMarketIterator<double> (unsupported/Eigen/src/SparseExtra/MatrixMarketIterator.h:41-241): walks a folder
    of MatrixMarket files.  ``matname.mtx`` is a matrix, ``matname_b.mtx`` its right-hand side, ``matname_x.mtx`` a
    reference solution; a symmetric file whose name contains "SPD" is flagged SPD (:222-224).  Usage as in the
    reference::

        it = MatrixMarketIterator(folder)
        while it:
            A, b = it.matrix(), it.rhs()
            ...
            it.next()
    """

    def __init__(self, folder: str):
        import os
        self._folder = folder
        self._valid_folder = os.path.isdir(folder)
        # readdir order is unspecified in the reference; sorted here so that runs are reproducible
        self._entries = sorted(os.listdir(folder)) if self._valid_folder else []
        self._pos = 0
        self._reset()
        self._isvalid = False
        if self._valid_folder:
            self._next_valid()

    def _reset(self):
        self._mat = None
        self._rhs = None
        self._refx = None
        self._has_rhs = self._has_refx = False

    def _next_valid(self):  # Getnextvalidmatrix, :193-227
        import os
        self._isvalid = False
        while self._pos < len(self._entries):
            name = self._entries[self._pos]
            self._pos += 1
            path = os.path.join(self._folder, name)
            if os.path.isdir(path):
                continue
            ok, sym, iscomplex, isvector = getMarketHeader(path)
            if not ok or isvector or iscomplex:   # Scalar = double: complex files are skipped (:205-214)
                continue
            if not name.endswith(".mtx"):
                continue
            self._matname = name[:-4]
            self._sym = SPD if ("SPD" in self._matname and sym != NonSymmetric) else sym
            self._isvalid = True
            break

    def __bool__(self):
        return self._isvalid

    def next(self):
        """operator++ (:67-74)."""
        self._reset()
        self._next_valid()
        return self

    def matname(self) -> str:
        return self._matname

    def sym(self) -> int:
        return self._sym

    def isFolderValid(self) -> bool:
        return self._valid_folder

    def hasRhs(self) -> bool:
        return self._has_rhs

    def hasrefX(self) -> bool:
        return self._has_refx

    def matrix(self) -> CsrMatrix:
        """:78-107: the matrix of the current file; a symmetric file that stores one triangle is expanded to the full
        matrix (the test `lower_norm > diag_norm && upper_norm == diag_norm` of the reference, on the stored entries)."""
        import os
        if self._mat is not None:
            return self._mat
        A = loadMarket(os.path.join(self._folder, self._matname + ".mtx"))
        if self._sym != NonSymmetric and A.rows == A.cols:
            rowof = np.repeat(np.arange(A.rows), np.diff(A.rowptr))
            d = np.sqrt(np.sum(A.vals[A.colidx == rowof] ** 2))
            lo = np.sqrt(np.sum(A.vals[A.colidx <= rowof] ** 2))
            up = np.sqrt(np.sum(A.vals[A.colidx >= rowof] ** 2))
            if (lo > d and up == d) or (up > d and lo == d):
                import scipy.sparse as sp
                S = A.to_scipy()
                S = (S + S.T - sp.diags(S.diagonal())).tocsr()
                S.sort_indices()
                A = CsrMatrix(A.rows, A.cols, S.indptr.astype(np.int32), S.indices.astype(np.int32),
                              np.ascontiguousarray(S.data, np.float64), 0, A.name)
        self._mat = A
        return A

    def rhs(self) -> np.ndarray:
        """:112-133: matname_b.mtx if it exists, else b = A * refX for a random refX (uniform in [-1, 1], as setRandom)."""
        import os
        if self._has_rhs:
            return self._rhs
        path = os.path.join(self._folder, self._matname + "_b.mtx")
        if os.path.exists(path):
            try:
                self._rhs = loadMarketVector(path)
                self._has_rhs = True
            except (ValueError, OSError):
                self._has_rhs = False
        if not self._has_rhs:
            A = self.matrix()
            import zlib
            rng = np.random.default_rng(zlib.crc32(self._matname.encode()))  # reproducible per matrix name
            self._refx = rng.uniform(-1.0, 1.0, A.cols)
            self._rhs = np.asarray(A.to_scipy() @ self._refx)
            self._has_refx = True
            self._has_rhs = True
        return self._rhs

    def refX(self) -> np.ndarray:
        """:141-156: matname_x.mtx if it exists (or the random solution behind a generated rhs), else an empty vector."""
        import os
        if self._has_refx:
            return self._refx
        path = os.path.join(self._folder, self._matname + "_x.mtx")
        if os.path.exists(path):
            try:
                self._refx = loadMarketVector(path)
                self._has_refx = True
            except (ValueError, OSError):
                self._has_refx = False
        if not self._has_refx:
            self._refx = np.zeros(0)
        return self._refx

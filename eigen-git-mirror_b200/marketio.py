"""MatrixMarket I/O with the semantics of the reference's SparseExtra module
(/root/reference/unsupported/Eigen/src/SparseExtra/MarketIO.h:109-282), the on-disk format of the reference's
real-matrix test flow (test/sparse_solver.h:147-179, bench/spbench/spbenchsolver.h:213-300).  SURVEY.md 8f rank 4.

* ``getMarketHeader``  -> (sym, iscomplex, isvector); sym uses Eigen's values Symmetric = 0x10, SelfAdjoint = 0x20
                          (Core/util/Constants.h), 0 for general.
* ``loadMarket``       -> CsrMatrix with the entries EXACTLY as stored: a symmetric file yields one triangle (the
                          caller picks UpLo, as spbenchsolver does), duplicates are summed (setFromTriplets),
                          out-of-range entries are skipped with a warning.
* ``saveMarket`` / ``loadMarketVector`` / ``saveMarketVector`` -- coordinate / array real, 17 significant digits.

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

"""Synthetic CSR workloads of BASELINE.json, built directly in CSR with numpy.

All matrices are compressed row-major CSR with sorted int32 column indices -- the arrays an
``Eigen::SparseMatrix<double,RowMajor,int>`` exposes through ``outerIndexPtr/innerIndexPtr/valuePtr``
(/root/reference/Eigen/src/SparseCore/SparseMatrix.h:149-183).  Grid ordering is ``row = i + n*j (+ n*n*k)`` as in
/root/reference/doc/special_examples/Tutorial_sparse_example_details.cpp:8-35; Dirichlet boundaries drop the
out-of-grid neighbours.

Every stencil generator takes an optional ``rows=(r0, r1)`` range so that a rank of a row-partitioned run builds only
its own block (column indices stay GLOBAL); nothing here ever materialises triplets.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

__all__ = [
    "CsrMatrix", "poisson2d", "poisson3d", "convdiff3d", "stencil27", "banded", "powerlaw", "rhs_from_solution",
    "random_vector", "varcoef3d",
]


@dataclass
class CsrMatrix:
    """A block of rows [row0, row0+rows) of a (global_rows x cols) CSR matrix; rowptr is local (starts at 0)."""
    rows: int
    cols: int
    rowptr: np.ndarray   # int32 [rows+1]
    colidx: np.ndarray   # int32 [nnz]
    vals: np.ndarray     # float64 / float32 [nnz]
    row0: int = 0
    name: str = ""

    @property
    def nnz(self) -> int:
        return int(self.rowptr[-1])

    def astype(self, dtype) -> "CsrMatrix":
        return CsrMatrix(self.rows, self.cols, self.rowptr, self.colidx, self.vals.astype(dtype), self.row0, self.name)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.vals, self.colidx, self.rowptr), shape=(self.rows, self.cols))

    def spmv_bytes(self) -> int:
        """Algorithmic bytes of y = A x: nnz*(S+4) + (rows+1)*4 + cols*S + rows*S."""
        s = self.vals.dtype.itemsize
        return self.nnz * (s + 4) + (self.rows + 1) * 4 + self.cols * s + self.rows * s


def _stencil(dims, offsets_vals, rows, dtype, chunk=1 << 22):
    """Generic constant-coefficient stencil on a tensor grid with natural ordering.

    ``offsets_vals`` is a list of ((di, dj, dk), value) sorted by resulting column offset."""
    dims = tuple(int(d) for d in dims) + (1,) * (3 - len(dims))
    nx, ny, nz = dims
    N = nx * ny * nz
    r0, r1 = (0, N) if rows is None else (int(rows[0]), int(rows[1]))
    assert 0 <= r0 <= r1 <= N
    if N >= 2 ** 31:
        raise ValueError("grid too large for int32 column indices")
    lin = [di + nx * dj + nx * ny * dk for (di, dj, dk), _ in offsets_vals]
    assert lin == sorted(lin), "stencil offsets must be sorted by column"
    counts = np.empty(r1 - r0, dtype=np.int32)
    col_chunks, val_chunks = [], []
    svals = np.array([v for _, v in offsets_vals], dtype=dtype)
    for c0 in range(r0, r1, chunk):
        c1 = min(c0 + chunk, r1)
        row = np.arange(c0, c1, dtype=np.int64)
        i = row % nx
        j = (row // nx) % ny
        k = row // (nx * ny)
        valid = np.empty((c1 - c0, len(offsets_vals)), dtype=bool)
        for s, ((di, dj, dk), _) in enumerate(offsets_vals):
            valid[:, s] = ((i + di >= 0) & (i + di < nx) & (j + dj >= 0) & (j + dj < ny)
                           & (k + dk >= 0) & (k + dk < nz))
        cols = (row[:, None] + np.array(lin, dtype=np.int64)[None, :]).astype(np.int32)
        counts[c0 - r0:c1 - r0] = valid.sum(axis=1, dtype=np.int32)
        col_chunks.append(cols[valid])
        val_chunks.append(np.broadcast_to(svals, valid.shape)[valid])
    rowptr = np.zeros(r1 - r0 + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    assert rowptr[-1] < 2 ** 31
    colidx = np.concatenate(col_chunks) if col_chunks else np.zeros(0, np.int32)
    vals = np.concatenate(val_chunks) if val_chunks else np.zeros(0, dtype)
    return CsrMatrix(r1 - r0, N, rowptr.astype(np.int32), np.ascontiguousarray(colidx),
                     np.ascontiguousarray(vals), r0)


def poisson2d(n: int, rows=None, dtype=np.float64) -> CsrMatrix:
    """2D 5-point Poisson on n x n, diag 4, off-diag -1 (config 0 of BASELINE.json at n=1024)."""
    st = [((0, -1, 0), -1.0), ((-1, 0, 0), -1.0), ((0, 0, 0), 4.0), ((1, 0, 0), -1.0), ((0, 1, 0), -1.0)]
    m = _stencil((n, n), st, rows, dtype)
    m.name = f"poisson2d_{n}"
    return m


def poisson3d(n: int, rows=None, dtype=np.float64) -> CsrMatrix:
    """3D 7-point Poisson on n^3, diag 6, off-diag -1 (configs 1 and 4 at n=256 / 512)."""
    st = [((0, 0, -1), -1.0), ((0, -1, 0), -1.0), ((-1, 0, 0), -1.0), ((0, 0, 0), 6.0),
          ((1, 0, 0), -1.0), ((0, 1, 0), -1.0), ((0, 0, 1), -1.0)]
    m = _stencil((n, n, n), st, rows, dtype)
    m.name = f"poisson3d_{n}"
    return m


def convdiff3d(n: int, gamma: float = 0.5, rows=None, dtype=np.float64) -> CsrMatrix:
    """3D convection-diffusion, central differences: lower neighbours -1-gamma, upper -1+gamma, diag 6 (config 2)."""
    lo, up = -1.0 - gamma, -1.0 + gamma
    st = [((0, 0, -1), lo), ((0, -1, 0), lo), ((-1, 0, 0), lo), ((0, 0, 0), 6.0),
          ((1, 0, 0), up), ((0, 1, 0), up), ((0, 0, 1), up)]
    m = _stencil((n, n, n), st, rows, dtype)
    m.name = f"convdiff3d_{n}_g{gamma}"
    return m


def stencil27(n: int, rows=None, dtype=np.float64) -> CsrMatrix:
    """27-point stencil on n^3: diag 26, all 26 neighbours -1 (SpMV sweep, config 3)."""
    st = []
    for dk in (-1, 0, 1):
        for dj in (-1, 0, 1):
            for di in (-1, 0, 1):
                st.append(((di, dj, dk), 26.0 if (di, dj, dk) == (0, 0, 0) else -1.0))
    m = _stencil((n, n, n), st, rows, dtype)
    m.name = f"stencil27_{n}"
    return m


def varcoef3d(n: int, seed: int = 7, dtype=np.float64) -> CsrMatrix:
    """Variable-coefficient symmetric 7-point operator (stresses rounding, unlike the constant
    Poisson matrix whose products are exact).  Edge weight w_e = 1 + 0.5*u_e, u_e ~ U(0,1) drawn per (lower
    endpoint, axis) so that A is symmetric; off-diagonals -w_e, diagonal = sum of incident weights + 0.25
    (strictly diagonally dominant, hence SPD)."""
    base = poisson3d(n, dtype=np.float64)
    rng = np.random.default_rng(seed)
    N = base.rows
    rowof = np.repeat(np.arange(N, dtype=np.int64), np.diff(base.rowptr))
    col = base.colidx.astype(np.int64)
    lo, dist = np.minimum(rowof, col), np.abs(rowof - col)
    axis = np.where(dist == 1, 0, np.where(dist == n, 1, 2))
    offdiag = dist != 0
    w_axis = 1.0 + 0.5 * rng.random((3, N))
    w = np.zeros(col.shape[0])
    w[offdiag] = w_axis[axis[offdiag], lo[offdiag]]
    rowsum = np.bincount(rowof[offdiag], weights=w[offdiag], minlength=N)
    vals = -w
    vals[~offdiag] = rowsum + 0.25
    return CsrMatrix(N, N, base.rowptr, base.colidx, vals.astype(dtype), 0, f"varcoef3d_{n}")


def banded(nrows: int, half_bw: int, seed: int = 12345, dtype=np.float64) -> CsrMatrix:
    """Banded matrix, nnz/row = 2k+1 away from the edges, values U(-1,1) (SpMV sweep)."""
    k = int(half_bw)
    row = np.arange(nrows, dtype=np.int64)
    lo = np.maximum(row - k, 0)
    hi = np.minimum(row + k, nrows - 1)
    counts = (hi - lo + 1)
    rowptr = np.zeros(nrows + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    nnz = int(rowptr[-1])
    assert nnz < 2 ** 31
    idx = np.arange(nnz, dtype=np.int64)
    rowof = np.repeat(row, counts)
    colidx = (lo[rowof] + (idx - rowptr[rowof])).astype(np.int32)
    rng = np.random.default_rng(seed)
    vals = rng.uniform(-1.0, 1.0, nnz).astype(dtype)
    return CsrMatrix(nrows, nrows, rowptr.astype(np.int32), colidx, vals, 0, f"banded_{nrows}_k{k}")


def powerlaw(nrows: int, mean_nnz: float, seed: int = 777, dtype=np.float64, max_row: int = 65536) -> CsrMatrix:
    """Power-law row lengths: nnz_i = clamp(floor(Pareto(alpha=2) scaled to mean m), 1, max_row); columns uniform
    random, sorted and distinct within each row (duplicates are dropped), values U(-1,1) (SpMV sweep)."""
    rng = np.random.default_rng(seed)
    raw = (rng.pareto(2.0, nrows) + 1.0)          # classical Pareto, x_m = 1, mean 2
    counts = np.clip(np.floor(raw * (mean_nnz / 2.0)), 1, min(max_row, nrows)).astype(np.int64)
    total = int(counts.sum())
    assert total < 2 ** 31
    rowof = np.repeat(np.arange(nrows, dtype=np.int64), counts)
    cols = rng.integers(0, nrows, total, dtype=np.int64)
    key = rowof * nrows + cols
    key.sort()
    keep = np.ones(total, dtype=bool)
    keep[1:] = key[1:] != key[:-1]
    key = key[keep]
    rowof = key // nrows
    colidx = (key - rowof * nrows).astype(np.int32)
    counts = np.bincount(rowof, minlength=nrows)
    rowptr = np.zeros(nrows + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    vals = rng.uniform(-1.0, 1.0, colidx.shape[0]).astype(dtype)
    return CsrMatrix(nrows, nrows, rowptr.astype(np.int32), colidx, vals, 0, f"powerlaw_{nrows}_m{mean_nnz:g}")


def random_vector(n: int, seed: int = 12345, dtype=np.float64) -> np.ndarray:
    """U(-1,1) vector (x_true of SURVEY.md 8d; numpy PCG64 stream instead of std::mt19937_64)."""
    return np.random.default_rng(seed).uniform(-1.0, 1.0, n).astype(dtype)


def rhs_from_solution(A: CsrMatrix, x_true: np.ndarray) -> np.ndarray:
    """b = A x_true for a full (row0 == 0, square) matrix, computed with scipy on the host (setup, untimed)."""
    return np.asarray(A.to_scipy() @ x_true)

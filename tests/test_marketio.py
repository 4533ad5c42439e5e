"""CPU: MatrixMarket I/O with the reference's semantics (unsupported/Eigen/src/SparseExtra/MarketIO.h)."""
import io
import os

import numpy as np
import pytest

from eigen_git_mirror_b200 import marketio as mio, workloads as wl


def test_roundtrip_matrix_and_vector(tmp_path):
    A = wl.varcoef3d(5)
    p = str(tmp_path / "a.mtx")
    assert mio.saveMarket(A, p)
    ok, sym, iscomplex, isvector = mio.getMarketHeader(p)
    assert ok and sym == 0 and not iscomplex and not isvector
    B = mio.loadMarket(p)
    assert B.rows == A.rows and np.array_equal(B.rowptr, A.rowptr) and np.array_equal(B.colidx, A.colidx)
    assert np.array_equal(B.vals, A.vals)  # 17 significant digits round-trip doubles exactly
    v = wl.random_vector(37, 3)
    pv = str(tmp_path / "v.mtx")
    assert mio.saveMarketVector(v, pv)
    assert mio.getMarketHeader(pv)[3] is True
    assert np.array_equal(mio.loadMarketVector(pv), v)


def test_agrees_with_scipy_reader(tmp_path):
    import scipy.io
    import scipy.sparse as sp
    A = wl.powerlaw(200, 5, seed=9)
    p = str(tmp_path / "p.mtx")
    mio.saveMarket(A, p)
    S = sp.csr_matrix(scipy.io.mmread(p))
    assert (abs(S - A.to_scipy())).max() == 0


def test_symmetric_file_yields_one_triangle_and_the_flag(tmp_path):
    """loadMarket keeps entries as stored; the symmetry is reported by the header and chosen by the caller as UpLo."""
    import scipy.sparse as sp
    A = wl.poisson2d(6)
    L = sp.tril(A.to_scipy()).tocsr()
    Lc = wl.CsrMatrix(A.rows, A.cols, L.indptr.astype(np.int32), L.indices.astype(np.int32), L.data)
    p = str(tmp_path / "s.mtx")
    mio.saveMarket(Lc, p, sym=mio.Symmetric)
    assert mio.getMarketHeader(p)[1] == mio.Symmetric
    B = mio.loadMarket(p)
    assert B.nnz == L.nnz and np.all(B.colidx <= np.repeat(np.arange(B.rows), np.diff(B.rowptr)))


def test_duplicates_are_summed_and_bad_entries_skipped(tmp_path, capsys):
    p = tmp_path / "d.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real general\n% comment\n3 3 5\n1 1 1.5\n1 1 2.5\n3 2 -1\n"
                 "4 1 9\n2 3 7\n")
    B = mio.loadMarket(str(p))
    assert B.to_scipy().toarray().tolist() == [[4.0, 0, 0], [0, 0, 7.0], [0, -1.0, 0]]
    assert "Invalid read" in capsys.readouterr().err


def test_missing_file():
    assert mio.getMarketHeader("/nonexistent/x.mtx")[0] is False

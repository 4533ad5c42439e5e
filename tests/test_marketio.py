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


def _write_folder(tmp_path):
    import scipy.sparse as sp
    A = wl.poisson2d(7)
    L = sp.tril(A.to_scipy()).tocsr()
    Lc = wl.CsrMatrix(A.rows, A.cols, L.indptr.astype(np.int32), L.indices.astype(np.int32), L.data)
    mio.saveMarket(Lc, str(tmp_path / "lap_SPD.mtx"), sym=mio.Symmetric)         # one triangle, SPD by name
    C = wl.convdiff3d(4)
    mio.saveMarket(C, str(tmp_path / "cd.mtx"))                                   # general
    x = wl.random_vector(C.rows, 4)
    mio.saveMarketVector(np.asarray(C.to_scipy() @ x), str(tmp_path / "cd_b.mtx"))
    mio.saveMarketVector(x, str(tmp_path / "cd_x.mtx"))
    (tmp_path / "notes.txt").write_text("not a matrix\n")
    (tmp_path / "sub").mkdir()
    (tmp_path / "cplx.mtx").write_text("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1.0 0.0\n")
    return A, C, x


def test_matrix_market_iterator_walks_a_folder_like_the_reference(tmp_path):
    """SparseExtra/MatrixMarketIterator.h: vectors, complex files, directories and non-MatrixMarket files are skipped;
    name_b.mtx / name_x.mtx are picked up; a symmetric one-triangle file is expanded; "SPD" in the name flags SPD."""
    A, C, x = _write_folder(tmp_path)
    it = mio.MatrixMarketIterator(str(tmp_path))
    assert it.isFolderValid()
    seen = {}
    while it:
        M = it.matrix()
        b = it.rhs()
        seen[it.matname()] = (it.sym(), M, b, it.refX().copy(), it.hasRhs(), it.hasrefX())
        it.next()
    assert sorted(seen) == ["cd", "lap_SPD"]
    sym, M, b, rx, has_b, has_x = seen["cd"]
    assert sym == mio.NonSymmetric and np.array_equal(M.vals, C.vals) and has_b and has_x
    assert np.array_equal(rx, x) and np.array_equal(b, np.asarray(C.to_scipy() @ x))
    sym, M, b, rx, has_b, has_x = seen["lap_SPD"]
    assert sym == mio.SPD
    assert (abs(M.to_scipy() - A.to_scipy())).max() == 0, "the stored triangle must be expanded to the full matrix"
    assert has_b and has_x and np.allclose(M.to_scipy() @ rx, b)   # generated: b = A * refX with a random refX
    assert not mio.MatrixMarketIterator(str(tmp_path / "missing")).isFolderValid()


def test_solve_market_tool_logic_with_stub_solvers(tmp_path, capsys):
    """tools/solve_market.py (the spbenchsolver-style flow) with the device solvers replaced by a direct solve: the folder
    walk, the solver list per matrix kind, the RCM permutation handed to the real host factorizations, the reports and
    the solution files.  tests/test_gpu_precond.py runs the tool itself on a B200."""
    import sys
    import types
    import scipy.sparse.linalg as spla
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import solve_market as sm
    import eigen_git_mirror_b200 as real

    class Stub:
        def __init__(self, A, preconditioner=None, **kw):
            self.A = A
            if preconditioner is not None and not isinstance(preconditioner, int):
                preconditioner.compute(A)           # the real host factorization, with the tool's permutation
                assert preconditioner.info() == 0
        def info(self): return 0
        def setTolerance(self, t): pass
        def setMaxIterations(self, n): pass
        def solve(self, b): return spla.spsolve(self.A.to_scipy().tocsc(), b)
        def stats(self): return {"last_solve_ms": 1.0}
        def iterations(self): return 7
        def error(self): return 1e-11
        def close(self): pass

    egm = types.SimpleNamespace(ConjugateGradient=Stub, BiCGSTAB=Stub, GMRES=Stub, IncompleteCholesky=real.IncompleteCholesky,
                                IncompleteLUT=real.IncompleteLUT, Lower=1, Success=0, B200Error=real.B200Error)
    _write_folder(tmp_path)
    out = tmp_path / "out"
    out.mkdir()
    args = types.SimpleNamespace(solvers="cg,cg_ic,bicgstab,bicgstab_ilut,gmres_ilut".split(","), ordering="rcm", droptol=1e-3,
                                 fillfactor=10, tol=1e-10, maxit=-1, out=str(out))
    it = mio.MatrixMarketIterator(str(tmp_path))
    while it:
        sm.run_one(egm, it.matname(), it.matrix(), it.rhs(), it.refX(), it.sym(), args)
        it.next()
    text = capsys.readouterr().out
    assert "== cd:" in text and "== lap_SPD:" in text and "general" in text and "SPD/symmetric" in text
    files = sorted(os.listdir(out))
    assert "lap_SPD_cg_ic_x.mtx" in files and "cd_gmres_ilut_x.mtx" in files and "cd_cg_x.mtx" not in files
    perm = sm.ordering_perm(mio.MatrixMarketIterator(str(tmp_path)).matrix(), "rcm")
    assert np.array_equal(np.sort(perm), np.arange(perm.size)) and sm.ordering_perm(None, "natural") is None

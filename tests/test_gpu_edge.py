"""GPU: edge cases of the drop-in path -- the ones the reference's tests exercise (test/sparse_product.cpp:336-378 zero
and 1x1 products, test/sparse_solver.h:41-145 uncompressed / re-compute / multi-column right-hand sides) plus the
corners of this implementation (rows longer than a shared-memory tile, empty rows, empty matrix, API misuse)."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _csr(S, dtype=np.float64):
    from eigen_git_mirror_b200.workloads import CsrMatrix
    S = S.tocsr()
    S.sort_indices()
    return CsrMatrix(S.shape[0], S.shape[1], S.indptr.astype(np.int32), S.indices.astype(np.int32),
                     S.data.astype(dtype))


def _scaled_ok(A, x, y, yref, eps):
    import scipy.sparse as sp
    M = sp.csr_matrix((np.abs(A.vals.astype(np.float64)), A.colidx, A.rowptr), shape=(A.rows, A.cols))
    scale = M @ np.abs(x.astype(np.float64))
    return np.all(np.abs(y.astype(np.float64) - yref.astype(np.float64)) <= eps * scale + 1e-300)


def test_one_by_one_and_zero_matrix(egm, port):
    """bug_942-style 1x1 product and the zero-matrix product of the reference's sparse_product test."""
    from eigen_git_mirror_b200.workloads import CsrMatrix
    one = CsrMatrix(1, 1, np.array([0, 1], np.int32), np.array([0], np.int32), np.array([2.5]))
    assert egm.SparseOperator(one).multiply(np.array([4.0]))[0] == 10.0
    for S in (egm.ConjugateGradient(one), egm.BiCGSTAB(one)):
        x = S.solve(np.array([5.0]))
        assert abs(x[0] - 2.0) < 1e-15 and S.info() == egm.Success
    zero = CsrMatrix(7, 7, np.zeros(8, np.int32), np.zeros(0, np.int32), np.zeros(0))
    y = egm.SparseOperator(zero).multiply(np.arange(7.0))
    assert y.shape == (7,) and not y.any()
    assert np.array_equal(egm.SparseOperator(zero).invdiag(), np.ones(7))  # absent diagonal -> 1


def test_empty_matrix(egm):
    from eigen_git_mirror_b200.workloads import CsrMatrix
    E = CsrMatrix(0, 0, np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0))
    s = egm.ConjugateGradient(E)
    x = s.solve(np.zeros(0))
    assert x.shape == (0,) and s.iterations() == 0 and s.info() == egm.Success


@pytest.mark.parametrize("dtype,eps", [(np.float64, 1e-13), (np.float32, 4e-6)])
def test_long_rows_empty_rows_and_ragged_tiles(dtype, eps, egm, port):
    """Rows longer than one shared-memory tile (2048 nnz), empty rows, a dense row next to singletons."""
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    n = 9000
    S = sp.random(n, n, density=0.0008, random_state=rng, data_rvs=lambda k: rng.uniform(-1, 1, k)).tolil()
    S[17, :] = rng.uniform(-1, 1, n)              # 9000 nnz: long tile
    S[4000, ::2] = rng.uniform(-1, 1, n // 2)     # 4500 nnz: long tile
    S[4001, :3000] = 1.0                          # 3000 nnz
    for r in (0, 5, 18, 8999):
        S[r, :] = 0                               # empty rows, also first and last
    A = _csr(S.tocsr(), dtype)
    x = rng.uniform(-1, 1, n).astype(dtype)
    for impl in (1, 2):
        op = egm.SparseOperator(A, spmv_impl=impl)
        y = op.multiply(x)
        st = op.stats()
        assert _scaled_ok(A, x, y, port.spmv(A, x), eps)
        assert y[0] == 0 and y[5] == 0 and y[8999] == 0
        if impl == 1:
            assert st["tiles_long"] >= 3
        op.close()


def test_uncompressed_input(egm, port):
    """innerNonZeroPtr input (test/sparse_solver.h:125-132): slots past inner_nnz hold garbage and must be ignored."""
    from eigen_git_mirror_b200 import workloads as wl
    A = wl.varcoef3d(8)
    lens = np.diff(A.rowptr)
    pad = 3
    rowptr = np.zeros(A.rows + 1, np.int32)
    rowptr[1:] = np.cumsum(lens + pad)
    colidx = np.full(rowptr[-1], 12345678, np.int32)
    vals = np.full(rowptr[-1], np.nan)
    for i in range(A.rows):
        colidx[rowptr[i]:rowptr[i] + lens[i]] = A.colidx[A.rowptr[i]:A.rowptr[i + 1]]
        vals[rowptr[i]:rowptr[i] + lens[i]] = A.vals[A.rowptr[i]:A.rowptr[i + 1]]
    U = wl.CsrMatrix(A.rows, A.cols, rowptr, colidx, vals)
    b = wl.rhs_from_solution(A, wl.random_vector(A.rows, 1))
    s = egm.ConjugateGradient()
    s.compute(U, inner_nnz=lens.astype(np.int32))
    s.setTolerance(1e-10)
    x = s.solve(b)
    xr, itr, _, _ = port.cg(A, b, tol=1e-10)
    assert s.info() == egm.Success and abs(s.iterations() - itr) <= 1
    assert np.linalg.norm(x - xr) / np.linalg.norm(xr) < 1e-8


@pytest.mark.parametrize("uplo", [1, 2])
def test_half_stored_symmetric_matrix(uplo, egm, port):
    """ConjugateGradient<_, Lower> / <_, Upper> on a matrix that stores only that triangle."""
    import scipy.sparse as sp
    from eigen_git_mirror_b200 import workloads as wl
    A = wl.varcoef3d(9)
    S = A.to_scipy()
    H = _csr(sp.tril(S) if uplo == 1 else sp.triu(S))
    b = wl.rhs_from_solution(A, wl.random_vector(A.rows, 2))
    s = egm.ConjugateGradient(H, uplo=uplo)
    s.setTolerance(1e-10)
    x = s.solve(b)
    xr, itr, _, infor = port.cg(H, b, tol=1e-10, uplo=uplo)   # the oracle reads one triangle like selfadjointView
    assert s.info() == infor == 0 and abs(s.iterations() - itr) <= 1
    assert np.linalg.norm(x - xr) / np.linalg.norm(xr) < 1e-8
    assert port.true_residual(A, x, b) < 2e-10


def test_multi_column_rhs_and_reuse(egm, port):
    """solve(B): sequential per column, info = worst, iterations()/error() = last column (IterativeSolverBase.h:375-388);
    the same solver object is then re-computed on a matrix of another size."""
    from eigen_git_mirror_b200 import workloads as wl
    A = wl.poisson2d(20)
    rng = np.random.default_rng(5)
    B = rng.uniform(-1, 1, (A.rows, 3))
    s = egm.ConjugateGradient(A)
    s.setTolerance(1e-10)
    X = s.solve(B)
    last = port.cg(A, np.ascontiguousarray(B[:, 2]), tol=1e-10)
    assert s.iterations() == last[1] and s.info() == egm.Success
    for k in range(3):
        xr = port.cg(A, np.ascontiguousarray(B[:, k]), tol=1e-10)[0]
        assert np.linalg.norm(X[:, k] - xr) / np.linalg.norm(xr) < 1e-8
    s.setMaxIterations(2)               # every column stops early -> NoConvergence overall
    s.solve(B)
    assert s.info() == egm.NoConvergence and s.iterations() == 2
    A2 = wl.convdiff3d(7)
    s2 = egm.BiCGSTAB(A)
    s2.compute(A2)                      # re-compute with another size and pattern
    s2.setTolerance(1e-10)
    b2 = wl.rhs_from_solution(A2, wl.random_vector(A2.rows, 3))
    x2 = s2.solve(b2)
    xr2 = port.bicgstab(A2, b2, tol=1e-10)[0]
    assert s2.info() == egm.Success and np.linalg.norm(x2 - xr2) / np.linalg.norm(xr2) < 1e-8


def test_api_misuse_is_reported(egm):
    from eigen_git_mirror_b200 import workloads as wl
    A = wl.poisson2d(8)
    s = egm.ConjugateGradient()
    with pytest.raises(AssertionError):
        s.solve(np.ones(A.rows))        # not initialized (IterativeSolverBase.h:337)
    with pytest.raises(AssertionError):
        s.factorize(A)                  # analyzePattern first (:218)
    s.compute(A)
    with pytest.raises(AssertionError):
        s.solve(np.ones(A.rows + 1))    # wrong number of rows (:338)
    op = egm.SparseOperator()
    with pytest.raises(egm.B200Error):
        op._dtype = np.float64
        op._rows = op._cols = 4
        op.multiply(np.ones(4))         # C ABI refuses: no matrix yet


def test_identity_preconditioner_and_guess(egm, port):
    from eigen_git_mirror_b200 import workloads as wl
    A = wl.varcoef3d(8)
    b = wl.rhs_from_solution(A, wl.random_vector(A.rows, 4))
    s = egm.ConjugateGradient(A, preconditioner=egm.IdentityPreconditioner)
    s.setTolerance(1e-10)
    x = s.solve(b)
    xr, itr, _, _ = port.cg(A, b, tol=1e-10, precond=0)
    assert abs(s.iterations() - itr) <= 1 and np.linalg.norm(x - xr) / np.linalg.norm(xr) < 1e-8
    assert np.array_equal(s.invdiag(), np.ones(A.rows))
    x0 = x + 1e-4
    xg = s.solveWithGuess(b, x0)
    xgr, itg, _, _ = port.cg(A, b, x0=x0, tol=1e-10, precond=0)
    assert abs(s.iterations() - itg) <= 1 and np.linalg.norm(xg - xgr) / np.linalg.norm(xgr) < 1e-8

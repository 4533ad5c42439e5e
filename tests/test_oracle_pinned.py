"""CPU, build container only: live comparison of the C oracle with the unmodified reference (oracle/_ref) on fresh
random inputs, beyond the committed golden vectors.  Skipped where oracle/_ref cannot exist (no /root/reference and no
prebuilt library)."""
import numpy as np
import pytest

from oracle import loader
from eigen_git_mirror_b200 import workloads as wl

pytestmark = [pytest.mark.timeout(300),
              pytest.mark.skipif(not loader.ref_available(), reason="oracle/_ref not built (needs /root/reference)")]


@pytest.mark.parametrize("seed", range(4))
def test_random_rows_spmv(seed):
    P, R = loader.port(), loader.ref()
    A = wl.powerlaw(500, 6 + 5 * seed, seed=seed, max_row=200)
    for dt in (np.float64, np.float32):
        Ad = A.astype(dt)
        x = wl.random_vector(A.cols, seed + 100, dt)
        assert np.array_equal(P.spmv(Ad, x), R.spmv(Ad, x))
        assert np.array_equal(R.spmv(Ad, x, threads=1), R.spmv(Ad, x, threads=4))  # OpenMP rows are independent


@pytest.mark.parametrize("n", [1, 2, 7, 8, 15, 16, 17, 31, 33, 100, 1001])
def test_dot_matches_reference_reduction_order(n):
    """||r||^2 as the first thing a CG solve computes: error() after 0 iterations exposes squaredNorm bit-for-bit."""
    P, R = loader.port(), loader.ref()
    import scipy.sparse as sp
    A = sp.identity(n, format="csr")
    M = wl.CsrMatrix(n, n, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64))
    b = wl.random_vector(n, n)
    x0 = wl.random_vector(n, n + 1)
    _, it_r, err_r, _ = R.cg(M, b, x0=x0, tol=1e-30, max_iters=0)
    _, it_p, err_p, _ = P.cg(M, b, x0=x0, tol=1e-30, max_iters=0)
    assert (it_r, err_r) == (it_p, err_p)


def test_solvers_on_fresh_matrices():
    P, R = loader.port(), loader.ref()
    A = wl.varcoef3d(9, seed=3)
    b = wl.rhs_from_solution(A, wl.random_vector(A.rows, 5))
    for uplo in (3, 1, 2):
        xr, itr, er, ir = R.cg(A, b, tol=1e-9, uplo=uplo)
        xp, itp, ep, ip = P.cg(A, b, tol=1e-9, uplo=uplo)
        assert (itr, er, ir) == (itp, ep, ip) and np.array_equal(xr, xp)
    C = wl.convdiff3d(9, gamma=0.7)
    b = wl.rhs_from_solution(C, wl.random_vector(C.rows, 6))
    xr, itr, er, ir = R.bicgstab(C, b, tol=1e-9)
    xp, itp, ep, ip = P.bicgstab(C, b, tol=1e-9)
    assert (itr, er, ir) == (itp, ep, ip) and np.array_equal(xr, xp)

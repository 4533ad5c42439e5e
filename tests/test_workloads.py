"""CPU: the synthetic matrices of BASELINE.json have the documented shapes (SURVEY.md section 8 size table)."""
import numpy as np

from eigen_git_mirror_b200 import workloads as wl


def test_nnz_formulas():
    n = 16
    assert wl.poisson2d(n).nnz == 5 * n * n - 4 * n
    assert wl.poisson3d(n).nnz == 7 * n ** 3 - 6 * n * n
    assert wl.convdiff3d(n).nnz == 7 * n ** 3 - 6 * n * n
    assert wl.stencil27(n).nnz == (3 * n - 2) ** 3


def test_full_size_formulas_match_survey_table():
    assert 5 * 1024 ** 2 - 4 * 1024 == 5_238_784
    assert 7 * 256 ** 3 - 6 * 256 ** 2 == 117_047_296
    assert 7 * 512 ** 3 - 6 * 512 ** 2 == 937_951_232 < 2 ** 31


def test_sorted_symmetric_and_row_blocks():
    A = wl.poisson3d(6)
    S = A.to_scipy()
    assert (abs(S - S.T)).nnz == 0
    for i in range(A.rows):
        c = A.colidx[A.rowptr[i]:A.rowptr[i + 1]]
        assert np.all(np.diff(c) > 0)
    B = wl.poisson3d(6, rows=(36, 150))
    assert B.row0 == 36 and B.rows == 114
    lo, hi = A.rowptr[36], A.rowptr[150]
    assert np.array_equal(B.colidx, A.colidx[lo:hi]) and np.array_equal(B.vals, A.vals[lo:hi])
    assert np.array_equal(B.rowptr, A.rowptr[36:151] - lo)


def test_convdiff_is_nonsymmetric_and_varcoef_is_spd():
    C = wl.convdiff3d(5).to_scipy()
    assert (abs(C - C.T)).nnz > 0
    V = wl.varcoef3d(5).to_scipy()
    assert (abs(V - V.T)).max() == 0
    assert np.linalg.eigvalsh(V.toarray()).min() > 0


def test_powerlaw_and_banded():
    P = wl.powerlaw(2000, 8)
    lens = np.diff(P.rowptr)
    assert lens.min() >= 1 and 5 < lens.mean() < 12
    for i in (0, 17, 1999):
        c = P.colidx[P.rowptr[i]:P.rowptr[i + 1]]
        assert np.all(np.diff(c) > 0)
    B = wl.banded(100, 4)
    assert np.diff(B.rowptr).max() == 9 and np.diff(B.rowptr).min() == 5

"""CPU: host logic of analyze_pattern (csrc/plan.cpp) through the GPU-free b200s_plan_probe."""
import numpy as np
import pytest

from eigen_git_mirror_b200 import planning, workloads as wl


def test_stencil_tiles_cover_everything():
    A = wl.poisson3d(20)
    v = planning.probe(A)
    st = v.stats
    assert st["rows"] == A.rows and st["nnz"] == A.nnz and st["ghosts"] == 0
    assert st["tiles"] >= A.rows // 256
    # 7-point rows: one thread per row everywhere
    assert st["tiles_by_lanes"][0] >= st["tiles"] - 1 and st["tiles_stream"] == 0 and st["tiles_long"] == 0
    assert np.array_equal(v.local_colidx, A.colidx)


@pytest.mark.parametrize("k,expect_lg", [(4, 0), (16, 2), (50, 3), (100, 4)])
def test_lanes_follow_row_length(k, expect_lg):
    A = wl.banded(4096, k)
    st = planning.probe(A).stats
    lanes = st["tiles_by_lanes"]
    assert int(np.argmax(lanes)) == expect_lg, lanes


def test_long_rows_and_imbalanced_tiles():
    A = wl.powerlaw(3000, 8, max_row=3000)
    # graft one very long row
    import scipy.sparse as sp
    S = A.to_scipy().tolil()
    S[7, :] = 1.0
    S = S.tocsr()
    S.sort_indices()
    B = wl.CsrMatrix(3000, 3000, S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data)
    st = planning.probe(B).stats
    assert st["tiles_long"] == 1
    assert st["tiles_stream"] > 0  # strongly imbalanced tiles take the two-phase (balanced products) path
    assert sum(st["tiles_by_lanes"]) + st["tiles_stream"] + st["tiles_long"] == st["tiles"]


def test_tile_caps_are_configurable():
    A = wl.poisson2d(64)
    a = planning.probe(A, tile_nnz=512, tile_rows=64).stats
    b = planning.probe(A, tile_nnz=4096, tile_rows=512).stats
    assert a["tiles"] > b["tiles"]
    assert a["tiles"] >= A.rows // 64


def test_empty_rows_and_empty_matrix():
    rowptr = np.array([0, 0, 2, 2, 3], np.int32)
    A = wl.CsrMatrix(4, 4, rowptr, np.array([0, 3, 2], np.int32), np.array([1.0, 2.0, 3.0]))
    st = planning.probe(A).stats
    assert st["rows"] == 4 and st["nnz"] == 3 and st["tiles"] == 1
    E = wl.CsrMatrix(0, 0, np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0))
    assert planning.probe(E).stats["tiles"] == 0


def test_invalid_inputs_are_rejected():
    import eigen_git_mirror_b200 as egm
    A = wl.poisson2d(8)
    bad = wl.CsrMatrix(A.rows, A.cols, A.rowptr, A.colidx.copy(), A.vals)
    bad.colidx[3] = A.cols + 5
    with pytest.raises(egm.B200Error):
        planning.probe(bad)
    # rectangular matrices are fine for the product (the solvers reject them at solve time)
    rect = wl.CsrMatrix(3, 4, np.array([0, 1, 2, 3], np.int32), np.array([0, 1, 3], np.int32), np.ones(3))
    assert planning.probe(rect).stats["cols"] == 4


def _dense_from(rowptr, colidx, vals, n):
    import scipy.sparse as sp
    return sp.csr_matrix((vals, colidx, rowptr), shape=(n, n)).toarray()


@pytest.mark.parametrize("uplo", [1, 2])
def test_one_triangle_is_expanded_to_the_full_symmetric_matrix(uplo):
    """UpLo = Lower / Upper (ConjugateGradient.h:202-213): the device matrix is what selfadjointView<UpLo> denotes,
    whether the input stores only that triangle or the whole matrix (the other triangle is then ignored)."""
    import scipy.sparse as sp
    A = wl.varcoef3d(5)
    S = A.to_scipy()
    full = S.toarray()
    half = sp.tril(S) if uplo == 1 else sp.triu(S)
    junk = S.copy().tolil()
    junk_dense = full.copy()
    if uplo == 1:   # poison the triangle that must not be read
        junk_dense[np.triu_indices(A.rows, 1)] *= 7.0
    else:
        junk_dense[np.tril_indices(A.rows, -1)] *= 7.0
    J = sp.csr_matrix(junk_dense)
    for M in (half.tocsr(), J):
        M.sort_indices()
        In = wl.CsrMatrix(A.rows, A.rows, M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.copy())
        rowptr, colidx, src = planning.canonical_csr(In, uplo=uplo)
        got = _dense_from(rowptr, colidx, In.vals[src], A.rows)
        assert np.array_equal(got, full)
        for i in range(A.rows):  # rows stay sorted when the input rows are sorted
            assert np.all(np.diff(colidx[rowptr[i]:rowptr[i + 1]]) > 0)


def test_uncompressed_input_is_compressed():
    """innerNonZeroPtr semantics (SparseMatrix.h:176-183): only the first inner_nnz[i] slots of each row count."""
    A = wl.powerlaw(300, 6, seed=2)
    lens = np.diff(A.rowptr)
    pad = np.arange(A.rows) % 4
    rowptr = np.zeros(A.rows + 1, np.int32)
    rowptr[1:] = np.cumsum(lens + pad)
    colidx = np.full(rowptr[-1], 0, np.int32)
    vals = np.full(rowptr[-1], np.nan)
    for i in range(A.rows):
        colidx[rowptr[i]:rowptr[i] + lens[i]] = A.colidx[A.rowptr[i]:A.rowptr[i + 1]]
        vals[rowptr[i]:rowptr[i] + lens[i]] = A.vals[A.rowptr[i]:A.rowptr[i + 1]]
    U = wl.CsrMatrix(A.rows, A.cols, rowptr, colidx, vals)
    rp, ci, src = planning.canonical_csr(U, inner_nnz=lens.astype(np.int32))
    assert np.array_equal(rp, A.rowptr) and np.array_equal(ci, A.colidx) and np.array_equal(vals[src], A.vals)


def test_compressed_full_input_is_passed_through():
    A = wl.poisson2d(9)
    rp, ci, src = planning.canonical_csr(A)
    assert np.array_equal(rp, A.rowptr) and np.array_equal(ci, A.colidx) and np.array_equal(src, np.arange(A.nnz))


def _span(A, uplo=3, inner_nnz=None):
    from eigen_git_mirror_b200 import _lib
    from eigen_git_mirror_b200.solvers import _ptr
    inz = None if inner_nnz is None else np.ascontiguousarray(inner_nnz, np.int32)
    return _lib.lib().b200s_plan_probe_span(A.rows, int(A.colidx.shape[0]), _ptr(A.rowptr), _ptr(A.colidx), _ptr(inz), uplo)


@pytest.mark.parametrize("uplo", [3, 1])
def test_rebased_rowptr_reads_the_right_slots(uplo):
    """A Map / Ref of an inner panel: rowptr[0] = base > 0.  Entry k takes its value from slot base + k, and the
    number of value slots factorize stages must reach past the last of them (round-1 advisor finding)."""
    A = wl.poisson2d(6)
    base = 5
    rowptr = (A.rowptr + base).astype(np.int32)
    colidx = np.concatenate([np.full(base, 0, np.int32), A.colidx])
    vals = np.concatenate([np.full(base, np.nan), A.vals])
    R = wl.CsrMatrix(A.rows, A.cols, rowptr, colidx, vals)
    rp, ci, src = planning.canonical_csr(R, uplo=uplo)
    assert src.min() >= base and src.max() < _span(R, uplo) == base + A.nnz
    got = _dense_from(rp, ci, vals[src], A.rows)
    assert np.array_equal(got, A.to_scipy().toarray())
    # the binding may pass either the entry count or the slot count as nnz
    R2 = wl.CsrMatrix(A.rows, A.cols, rowptr, colidx[: base + A.nnz], vals)
    assert _span(R2, uplo) == base + A.nnz


def test_span_of_uncompressed_and_plain_inputs():
    A = wl.poisson2d(5)
    assert _span(A) == A.nnz
    lens = np.diff(A.rowptr)
    rowptr = np.zeros(A.rows + 1, np.int32)
    rowptr[1:] = np.cumsum(lens + 2)
    colidx = np.zeros(rowptr[-1], np.int32)
    for i in range(A.rows):
        colidx[rowptr[i]:rowptr[i] + lens[i]] = A.colidx[A.rowptr[i]:A.rowptr[i + 1]]
    U = wl.CsrMatrix(A.rows, A.cols, rowptr, colidx, np.zeros(rowptr[-1]))
    assert _span(U, inner_nnz=lens) == rowptr[-2] + lens[-1]
    bad = wl.CsrMatrix(A.rows, A.cols, A.rowptr, A.colidx[:-3], A.vals[:-3])
    with pytest.raises(Exception):
        planning.probe(bad)  # rowptr references slots past nnz


def test_indices_are_narrowed_to_int32_before_crossing_the_abi():
    """A CsrMatrix built with int64 indices (np.cumsum, scipy with 64-bit indices) must not be read as int32 garbage."""
    A = wl.poisson2d(7)
    A64 = wl.CsrMatrix(A.rows, A.cols, A.rowptr.astype(np.int64), A.colidx.astype(np.int64), A.vals)
    v = planning.probe(A64)
    assert v.stats["nnz"] == A.nnz and np.array_equal(v.local_colidx, A.colidx)
    big = wl.CsrMatrix(2, 2 ** 33, np.array([0, 1, 2], np.int64), np.array([0, 2 ** 32 + 5], np.int64), np.ones(2))
    with pytest.raises(ValueError):
        planning.probe(big)


def test_rectangular_matrices_are_accepted_for_the_product():
    import scipy.sparse as sp
    S = sp.random(30, 70, density=0.1, random_state=np.random.default_rng(3)).tocsr()
    S.sort_indices()
    v = planning.probe(S)
    assert v.stats["rows"] == 30 and v.stats["cols"] == 70 and v.stats["nnz"] == S.nnz


def test_partition_rows_balances_bytes_when_given_rowptr():
    from eigen_git_mirror_b200 import partition_rows
    A = wl.powerlaw(50000, 16, seed=9)
    even = partition_rows(A.rows, 8)
    bal = partition_rows(A.rows, 8, rowptr=A.rowptr)
    assert bal[0] == 0 and bal[-1] == A.rows and np.all(np.diff(bal) > 0)
    cost = lambda st: np.diff(12 * A.rowptr[st].astype(np.int64) + 104 * st)
    assert cost(bal).max() <= 1.02 * cost(bal).mean()
    assert cost(bal).max() <= cost(even).max()
    n = 16
    P = wl.poisson3d(n)
    assert np.array_equal(partition_rows(P.rows, 4, align=n * n, rowptr=P.rowptr), partition_rows(P.rows, 4, align=n * n))

"""CPU: incomplete-factorization preconditioners, host half.

1. the restated factorizations (csrc/factors.cpp) give the factors of the UNMODIFIED reference entry for entry and bit for
   bit: IncompleteLUT::m_lu for the reference's own AMD permutation, IncompleteCholesky::m_L / m_scale for the natural
   ordering and for the reference's AMD permutation, Lower and Upper;
2. the staged form the device runs (rows grouped into dependency levels, csrc/kernels_tri.cuh), emulated on the CPU by
   oracle/oracle.c in LEVEL order, reproduces IncompleteLUT::solve / IncompleteCholesky::solve of the reference bit for
   bit -- i.e. stage extraction, level analysis, gathers / scalings and the per-step rounding are what the GPU needs;
3. structural properties of the level schedule and the launch plan; error handling.
Needs oracle/_ref (built where /root/reference exists, shipped prebuilt otherwise)."""
import numpy as np
import pytest
import scipy.sparse as sp

from eigen_git_mirror_b200 import workloads as wl
from eigen_git_mirror_b200._lib import B200Error
from eigen_git_mirror_b200.preconditioners import IncompleteCholesky, IncompleteLUT
from oracle import loader

pytestmark = pytest.mark.skipif(not loader.ref_available(), reason="oracle/_ref not built")


def _random_square(n, density, seed, dominant=True, symmetric=False):
    rng = np.random.default_rng(seed)
    S = sp.random(n, n, density=density, random_state=rng, format="csr")
    S.data = rng.standard_normal(S.nnz)
    if symmetric:
        S = (S + S.T) * 0.5
    d = np.asarray(abs(S).sum(axis=1)).ravel() + 1.0 if dominant else rng.standard_normal(n)
    S = (S + sp.diags(d)).tocsr()
    S.sort_indices()
    return wl.CsrMatrix(n, n, S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data.copy())


def matrices():
    return [("poisson3d_7", wl.poisson3d(7)), ("convdiff3d_8", wl.convdiff3d(8)), ("poisson2d_24", wl.poisson2d(24)),
            ("varcoef3d_8", wl.varcoef3d(8)), ("random_300", _random_square(300, 0.03, 11)),
            ("random_sym_257", _random_square(257, 0.04, 12, symmetric=True))]


MATS = matrices()
IDS = [m[0] for m in MATS]


@pytest.fixture(scope="module")
def ref3():
    """The AVX2+FMA build of the reference: its scalar code paths are the ones the restatement is pinned to (the AVX-512
    build replaces sqrt by an rsqrt14-based approximation in m_scale.cwiseSqrt(), EIGEN_FAST_MATH)."""
    return loader.Ref("v3")


@pytest.fixture(scope="module", params=["v3", "v4"])
def ref_any(request):
    flags = loader._cpu_flags()
    if request.param == "v4" and "avx512f" not in flags:
        pytest.skip("host has no AVX-512")
    return loader.Ref(request.param)


# ---------------------------------------------------------------------------------------------------- 1. factors
@pytest.mark.parametrize("name,A", MATS, ids=IDS)
@pytest.mark.parametrize("droptol,fillfactor", [(-1.0, 0), (1e-3, 10), (1e-12, 2), (0.05, 40)])
def test_ilut_factor_is_the_references(ref_any, name, A, droptol, fillfactor):
    rp, ci, va, P, Pinv, info = ref_any.ilut(A, droptol, fillfactor)
    f = IncompleteLUT(A, droptol, fillfactor, perm=P)
    outer, inner, vals, _, perm = f.arrays()
    assert f.info() == info == 0
    assert np.array_equal(outer, rp) and np.array_equal(inner, ci), "pattern of m_lu (incl. the QuickSplit order)"
    assert np.array_equal(vals, va), f"values of m_lu differ in {(vals != va).sum()} places"
    assert np.array_equal(perm, P)


@pytest.mark.parametrize("name,A", MATS, ids=IDS)
@pytest.mark.parametrize("uplo", [1, 2])
@pytest.mark.parametrize("ordering", [0, 1])
def test_ichol_factor_is_the_references(ref3, name, A, uplo, ordering):
    cp, ri, lv, sc, perm, info = ref3.ichol(A, uplo, ordering)
    g = IncompleteCholesky(A, uplo=uplo, perm=(perm if ordering else None))
    outer, inner, vals, scale, p = g.arrays()
    assert g.info() == info
    if info != 0:
        return  # the reference gave up after its 10 shifts (random unsymmetric input): nothing to compare
    assert np.array_equal(outer, cp) and np.array_equal(inner, ri), "pattern of m_L"
    assert np.array_equal(scale, sc), "m_scale"
    assert np.array_equal(vals, lv), f"values of m_L differ in {(vals != lv).sum()} places"
    assert np.array_equal(p, perm)


def test_ichol_shift_retries_follow_the_reference(ref3):
    """A symmetric matrix with a negative diagonal entry and weak diagonals: the first attempts hit a non-positive
    pivot and the factorization restarts with a doubled shift (IncompleteCholesky.h:322-338)."""
    base = _random_square(120, 0.08, 5, symmetric=True)
    S = base.to_scipy().tolil()
    for i in range(0, 120, 9):
        S[i, i] = -0.5
    for i in range(1, 120, 7):
        S[i, i] = 0.05
    S = S.tocsr()
    S.sort_indices()
    A = wl.CsrMatrix(120, 120, S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data.copy())
    for ordering in (0, 1):
        cp, ri, lv, sc, perm, info = ref3.ichol(A, 1, ordering)
        g = IncompleteCholesky(A, uplo=1, perm=(perm if ordering else None))
        outer, inner, vals, scale, _ = g.arrays()
        assert g.info() == info
        if info == 0:
            assert np.array_equal(outer, cp) and np.array_equal(inner, ri) and np.array_equal(vals, lv)
            assert np.array_equal(scale, sc)


def test_ilut_zero_row_is_a_numerical_issue(ref3):
    A = wl.poisson2d(6)
    vals = A.vals.copy()
    vals[A.rowptr[5]:A.rowptr[6]] = 0.0
    Z = wl.CsrMatrix(A.rows, A.cols, A.rowptr, A.colidx, vals)
    *_, info = ref3.ilut(Z)
    f = IncompleteLUT(Z)   # natural ordering: the zero row is hit whatever the permutation
    assert f.info() == info == 1  # NumericalIssue (IncompleteLUT.h:321-325)


# ----------------------------------------------------------------------------------- 2. the staged apply (CPU emulation)
@pytest.mark.parametrize("name,A", MATS, ids=IDS)
def test_staged_ilut_apply_has_the_references_bits(ref_any, port, name, A):
    r = wl.random_vector(A.rows, 7)
    *_, P, _, info = ref_any.ilut(A)
    f = IncompleteLUT(A, perm=P)
    want = ref_any.ilut_solve(A, r)
    got = port.factors_apply(f, r, order="level")
    assert np.array_equal(got, want), np.abs(got - want).max()
    assert np.array_equal(port.factors_apply(f, r, order="natural"), want)
    # the rounding of each step matters and is pinned: with FMAs the result differs
    assert not np.array_equal(port.factors_apply(f, r, order="level", fused=True), want)


@pytest.mark.parametrize("name,A", MATS[:4] + MATS[5:], ids=IDS[:4] + IDS[5:])
@pytest.mark.parametrize("uplo", [1, 2])
@pytest.mark.parametrize("ordering", [0, 1])
def test_staged_ichol_apply_has_the_references_bits(ref_any, port, name, A, uplo, ordering):
    """With the reference's own factor (whose m_scale depends on the ISA variant) the staged apply is bit-exact on every
    variant; the factor itself is pinned by test_ichol_factor_is_the_references."""
    r = wl.random_vector(A.rows, 8)
    cp, ri, lv, sc, perm, info = ref_any.ichol(A, uplo, ordering)
    assert info == 0
    g = IncompleteCholesky.from_factors(cp, ri, lv, sc, perm)
    want = ref_any.ichol_solve(A, r, uplo, ordering)
    got = port.factors_apply(g, r, order="level")
    assert np.array_equal(got, want), np.abs(got - want).max()
    assert [g.stage(0).fused, g.stage(1).fused] == [True, False]


def test_own_ichol_factor_end_to_end(ref3, port):
    A = wl.poisson3d(9)
    r = wl.random_vector(A.rows, 9)
    for ordering in (0, 1):
        *_, perm, info = ref3.ichol(A, 1, ordering)
        g = IncompleteCholesky(A, uplo=1, perm=(perm if ordering else None))
        assert np.array_equal(port.factors_apply(g, r), ref3.ichol_solve(A, r, 1, ordering))


# ------------------------------------------------------------------------------------------ 3. schedule and launch plan
@pytest.mark.parametrize("kind", ["ilut", "ichol"])
def test_level_schedule_and_launch_plan(kind):
    A = wl.poisson3d(14)  # 2,744 rows: levels wider than one CTA do not occur, so also a 2-D case below
    B = wl.poisson2d(90)  # 8,100 rows, natural ordering: anti-diagonal wavefronts up to 90 rows
    C_ = _random_square(6000, 0.0007, 3)  # sparse random: a few very wide levels (> 1024 rows)
    for M in (A, B, C_):
        f = IncompleteLUT(M) if kind == "ilut" else IncompleteCholesky(M, uplo=1)
        assert f.info() == 0
        n = M.rows
        for which in (0, 1):
            st = f.stage(which)
            nlev = len(st.level_ptr) - 1
            assert st.level_ptr[0] == 0 and st.level_ptr[-1] == n and np.all(np.diff(st.level_ptr) > 0)
            assert np.array_equal(np.sort(st.level_rows), np.arange(n))
            level_of = np.empty(n, np.int64)
            for l in range(nlev):
                level_of[st.level_rows[st.level_ptr[l]:st.level_ptr[l + 1]]] = l
            rows = np.repeat(np.arange(n), np.diff(st.rowptr))
            assert np.all(level_of[st.colidx] < level_of[rows]), "a row reads a row of the same or a later level"
            # forward stage reads smaller rows, backward stage larger ones
            assert np.all(st.colidx < rows) if which == 0 else np.all(st.colidx > rows)
            # every row sits exactly one level above its deepest dependency (levels are as shallow as possible)
            deepest = np.full(n, -1, np.int64)
            np.maximum.at(deepest, rows, level_of[st.colidx])
            assert np.array_equal(level_of, deepest + 1)
            # launches tile the levels in order; fused runs only hold levels of at most 1024 rows
            la = st.launches
            assert la[0, 0] == 0 and la[-1, 1] == nlev and np.array_equal(la[1:, 0], la[:-1, 1])
            widths = np.diff(st.level_ptr)
            for b, e, w in la:
                assert w == widths[b:e].max()
                if e - b > 1 or w <= 1024:
                    assert widths[b:e].max() <= 1024
                else:
                    assert e - b == 1 and w > 1024
    assert any(w > 1024 for w in f.stage(0).launches[:, 2]), "the random case should have a grid-wide level"


def test_factors_from_caller_arrays_equal_own_factorization(ref3):
    A = wl.convdiff3d(7)
    rp, ci, va, P, _, _ = ref3.ilut(A)
    a, b = IncompleteLUT.from_factors(rp, ci, va, P), IncompleteLUT(A, perm=P)
    for which in (0, 1):
        sa, sb = a.stage(which), b.stage(which)
        for x, y in zip((sa.rowptr, sa.colidx, sa.vals, sa.level_ptr, sa.level_rows, sa.launches),
                        (sb.rowptr, sb.colidx, sb.vals, sb.level_ptr, sb.level_rows, sb.launches)):
            assert np.array_equal(x, y)
    assert all(np.array_equal(x, y) for x, y in zip(a.permscale()[::2], b.permscale()[::2]))


def test_argument_errors():
    A = wl.poisson2d(5)
    with pytest.raises(B200Error, match="not a permutation"):
        IncompleteLUT(A, perm=np.zeros(A.rows, np.int32))
    with pytest.raises(B200Error, match="uplo"):
        IncompleteCholesky(A, uplo=3)
    S = A.to_scipy().tolil()
    S[3, 3] = 0.0
    S = S.tocsr()
    S.eliminate_zeros()
    nodiag = wl.CsrMatrix(A.rows, A.cols, S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data.copy())
    with pytest.raises(B200Error, match="diagonal"):
        IncompleteCholesky(nodiag, uplo=1)
    with pytest.raises(B200Error, match="diagonal"):
        IncompleteLUT.from_factors(np.array([0, 1, 2], np.int32), np.array([1, 0], np.int32), np.array([1.0, 1.0]))
    with pytest.raises(AssertionError):
        IncompleteLUT().info()


def test_multicolor_ordering_gives_few_wide_levels():
    """b200s_ordering_multicolor: a valid permutation (reference convention) that groups rows by colour; no two coupled
    rows share a colour; red-black on the 7-point stencil, so the incomplete-Cholesky solves have 2 levels instead of ~3n."""
    from eigen_git_mirror_b200.preconditioners import multicolor_ordering
    for A, want in ((wl.poisson3d(12), 2), (wl.poisson2d(17), 2), (_random_square(400, 0.02, 21, symmetric=True), None)):
        perm, nc = multicolor_ordering(A)
        assert np.array_equal(np.sort(perm), np.arange(A.rows))
        if want:
            assert nc == want
        S = A.to_scipy()
        S = (abs(S) + abs(S.T)).tocoo()
        off = S.row != S.col
        # colour of a vertex = the colour block its new position falls into: recover block boundaries from the ordering
        order = np.argsort(perm)                      # order[k] = old row at new position k
        P = sp.csr_matrix((np.ones(A.rows), (perm, np.arange(A.rows))), shape=(A.rows, A.rows))
        B = (P @ (abs(A.to_scipy()) + abs(A.to_scipy().T)) @ P.T).tocsr()
        B.setdiag(0)
        B.eliminate_zeros()
        # rows of one colour are contiguous and mutually uncoupled: greedy level analysis of the lower part of B (every
        # row one level above its deepest coupled predecessor) must give exactly nc levels
        level = np.zeros(A.rows, np.int64)
        for i in range(A.rows):
            cols = B.indices[B.indptr[i]:B.indptr[i + 1]]
            cols = cols[cols < i]
            if cols.size:
                level[i] = level[cols].max() + 1
        assert level.max() + 1 <= nc and np.all(np.diff(level) >= 0), "colour blocks must be contiguous and uncoupled"
        assert off.any()
    A = wl.poisson3d(12)
    perm, _ = multicolor_ordering(A)
    g = IncompleteCholesky(A, uplo=1, perm=perm)
    assert g.info() == 0 and [len(g.stage(w).level_ptr) - 1 for w in (0, 1)] == [2, 2]
    nat = IncompleteCholesky(A, uplo=1)
    assert len(nat.stage(0).level_ptr) - 1 == 3 * 12 - 2


def test_random_matrices_fuzz(ref3, port):
    """40 random systems (sizes 2..120, densities 2-40 %, explicit zeros, random droptol / fillfactor): factors and staged
    applies bit for bit as the reference's, for ILUT (reference's AMD permutation) and IC (Lower and Upper, AMD)."""
    rng = np.random.default_rng(0)

    def mk(S):
        S = S.tocsr()
        S.sort_indices()
        return wl.CsrMatrix(S.shape[0], S.shape[1], S.indptr.astype(np.int32), S.indices.astype(np.int32),
                            S.data.astype(np.float64))

    for t in range(40):
        n = int(rng.integers(2, 120))
        S = sp.random(n, n, density=float(rng.uniform(0.02, 0.4)), random_state=rng, format="csr")
        S.data = rng.standard_normal(S.nnz)
        if t % 3 == 0 and S.nnz:
            S.data[rng.integers(0, S.nnz, size=max(1, S.nnz // 10))] = 0.0   # stored zeros stay entries, as in Eigen
        S = S + sp.diags(np.asarray(abs(S).sum(axis=1)).ravel() + rng.uniform(0.1, 2, n))
        A = mk(S)
        dt, ff = (-1.0, 0) if t % 2 else (float(10 ** rng.uniform(-6, -1)), int(rng.integers(1, 30)))
        rp, ci, va, P, _, info = ref3.ilut(A, dt, ff)
        f = IncompleteLUT(A, dt, ff, perm=P)
        outer, inner, vals, _, _ = f.arrays()
        assert f.info() == info and np.array_equal(outer, rp) and np.array_equal(inner, ci) and np.array_equal(vals, va), t
        r = rng.standard_normal(n)
        if info == 0:
            assert np.array_equal(port.factors_apply(f, r), ref3.ilut_solve(A, r, dt, ff)), t
        Asym = mk((S + S.T) * 0.5)
        for uplo in (1, 2):
            cp, ri, lv, sc, perm, info = ref3.ichol(Asym, uplo, 1)
            g = IncompleteCholesky(Asym, uplo=uplo, perm=perm)
            assert g.info() == info, t
            if info == 0:
                outer, inner, vals, scale, _ = g.arrays()
                assert np.array_equal(outer, cp) and np.array_equal(inner, ri) and np.array_equal(vals, lv), t
                assert np.array_equal(scale, sc) and np.array_equal(port.factors_apply(g, r), ref3.ichol_solve(Asym, r, uplo, 1)), t


def test_degenerate_sizes(ref3, port):
    for dense in ([[2.5]], [[4.0, 1.0], [1.0, 3.0]]):
        S = sp.csr_matrix(np.array(dense))
        A = wl.CsrMatrix(S.shape[0], S.shape[1], S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data.copy())
        r = np.arange(1, A.rows + 1, dtype=np.float64)
        *_, P, _, _ = ref3.ilut(A)
        assert np.array_equal(port.factors_apply(IncompleteLUT(A, perm=P), r), ref3.ilut_solve(A, r))
        assert np.array_equal(port.factors_apply(IncompleteCholesky(A, uplo=1), r), ref3.ichol_solve(A, r, 1, 0))
    E = wl.CsrMatrix(0, 0, np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0))
    assert IncompleteLUT(E).info() == 0 and IncompleteCholesky(E, uplo=1).info() == 0


# ------------------------------------------------------ 4. the CPU restatement of the preconditioned solver loops
@pytest.mark.parametrize("name,A", [MATS[0], MATS[2], MATS[3], MATS[5]], ids=[IDS[0], IDS[2], IDS[3], IDS[5]])
@pytest.mark.parametrize("uplo,ordering", [(1, 0), (1, 1), (2, 1), (3, 1)])
def test_port_cg_with_incomplete_cholesky_is_the_references(ref_any, port, name, A, uplo, ordering):
    """oracle_cg_precond (oracle_body.h) + the staged IncompleteCholesky apply (oracle.c) against the unmodified
    ConjugateGradient<_, UpLo, IncompleteCholesky<double, UpLo or Lower, Natural | AMD>>: x, iterations(), error() and
    info() bit for bit, on both ISA builds (the reduction pattern follows the build, the factor is the build's own)."""
    lanes = 8 if ref_any.variant == "v4" else 4
    b = np.asarray(A.to_scipy() @ wl.random_vector(A.rows, 12345))
    cp, ri, lv, sc, perm, info = ref_any.ichol(A, 1 if uplo == 3 else uplo, ordering)
    assert info == 0
    pre = IncompleteCholesky.from_factors(cp, ri, lv, sc, perm)
    for mi in (-1, 3):
        want = ref_any.precond_solver("cg_ichol", A, b, tol=1e-10, max_iters=mi, uplo=uplo, ordering=ordering)
        got = port.cg_factors(A, b, pre, tol=1e-10, max_iters=mi, uplo=uplo, lanes=lanes)
        assert got[1:] == want[1:], (got[1:], want[1:])
        assert np.array_equal(got[0], want[0])


@pytest.mark.parametrize("name,A", [MATS[0], MATS[1], MATS[4]], ids=[IDS[0], IDS[1], IDS[4]])
def test_port_bicgstab_with_ilut_is_the_references(ref_any, port, name, A):
    lanes = 8 if ref_any.variant == "v4" else 4
    b = np.asarray(A.to_scipy() @ wl.random_vector(A.rows, 12345))
    for droptol, fill in ((-1.0, 0), (1e-2, 5)):
        rp, ci, va, P, _, info = ref_any.ilut(A, droptol, fill)
        pre = IncompleteLUT.from_factors(rp, ci, va, P)
        for mi in (-1, 2):
            want = ref_any.precond_solver("bicgstab_ilut", A, b, tol=1e-10, max_iters=mi, droptol=droptol, fillfactor=fill)
            got = port.bicgstab_factors(A, b, pre, tol=1e-10, max_iters=mi, lanes=lanes)
            assert (got[1], got[3]) == (want[1], want[3]), (got[1:], want[1:])
            if A.rows % 8 == 0:
                assert got[2] == want[2] and np.array_equal(got[0], want[0])
            else:
                # outside the BiCGSTAB port's pinned range (the scalar remainder of
                # `x += alpha*y + w*z` is contracted differently when the length is not a multiple of the packet)
                assert abs(got[2] - want[2]) <= 1e-9 * want[2]
                assert np.allclose(got[0], want[0], rtol=0, atol=1e-13 * np.abs(want[0]).max())

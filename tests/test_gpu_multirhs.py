"""GPU: multi-column right-hand sides.  The reference solves the columns of B one after the other
(IterativeSolverBase.h:366-389); the batched CG shares one stream of the matrix between up to 8 columns per iteration
and must return, for EVERY column, bit for bit what the single-column solve returns -- x, iterations(), error(),
info() -- with the aggregate iterations()/error()/info() of the reference's loop (last column / worst column)."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
TOL = 1e-10


def _columns(wl, A, ncols, seed=7):
    """Right-hand sides that converge at different iterations, plus the special cases of the control flow."""
    rng = np.random.default_rng(seed)
    S = A.to_scipy()
    cols = []
    for k in range(ncols):
        kind = k % 6
        if kind == 0:
            cols.append(np.asarray(S @ wl.random_vector(A.rows, 100 + k)))
        elif kind == 1:
            cols.append(np.ones(A.rows))
        elif kind == 2:
            cols.append(np.zeros(A.rows))                      # ||b|| = 0: x = 0, iterations 0, error 0
        elif kind == 3:
            cols.append(1e-3 * rng.standard_normal(A.rows))
        elif kind == 4:
            e = np.zeros(A.rows); e[A.rows // 3] = 1.0
            cols.append(e)
        else:
            cols.append(np.asarray(S @ np.sin(np.arange(A.rows) * 0.01)) * 1e6)
    return np.asfortranarray(np.stack(cols, axis=1))


def _sequential(s, B, X0=None):
    out = []
    for k in range(B.shape[1]):
        x = s.solve(B[:, k].copy()) if X0 is None else s.solveWithGuess(B[:, k].copy(), X0[:, k].copy())
        out.append((x, s.iterations(), s.error(), s.info()))
    return out


def _check_batched_equals_sequential(egm, s, B, X0=None):
    seq = _sequential(s, B, X0)
    X = s.solve(B) if X0 is None else s.solveWithGuess(B, X0)
    assert X.shape == B.shape
    for k, (x, it, err, info) in enumerate(seq):
        assert np.array_equal(X[:, k], x, equal_nan=True), (k, np.abs(X[:, k] - x).max())
        assert s.column_iterations[k] == it and s.column_infos[k] == info, (k, s.column_iterations[k], it)
        assert s.column_errors[k] == err or (np.isnan(err) and np.isnan(s.column_errors[k]))
    # aggregate state as the reference's loop leaves it (IterativeSolverBase.h:375-388)
    assert s.iterations() == seq[-1][1] and (s.error() == seq[-1][2] or np.isnan(seq[-1][2]))
    want = egm.Success
    for _, _, _, info in seq:
        if info == egm.NumericalIssue:
            want = egm.NumericalIssue
        elif info == egm.NoConvergence:
            want = egm.NoConvergence
    assert s.info() == want
    return seq


@pytest.mark.parametrize("ncols", [2, 3, 4, 5, 7, 8, 11, 17])
@pytest.mark.parametrize("kmax", [2, 4, 8])
def test_batched_columns_equal_single_column_solves(ncols, kmax, egm, monkeypatch):
    """Every batch width the library has kernels for (B200S_MULTI_K; 4 is the default, measured fastest)."""
    from eigen_git_mirror_b200 import workloads as wl
    monkeypatch.setenv("B200S_MULTI_K", str(kmax))
    A = wl.varcoef3d(14)
    s = egm.ConjugateGradient(A)
    assert s.multi_rhs_batch() == kmax
    s.setTolerance(TOL)
    seq = _check_batched_equals_sequential(egm, s, _columns(wl, A, ncols))
    assert len({it for _, it, _, _ in seq}) > 1  # the columns really stop at different iterations
    s.close()


@pytest.mark.parametrize("matrix", ["poisson2d", "banded16", "banded50", "stencil27", "random_spd"])
@pytest.mark.parametrize("loop_mode", [1, 2, 3, 4])
def test_batched_over_tile_flavours_and_loop_modes(matrix, loop_mode, egm, golden):
    from eigen_git_mirror_b200 import workloads as wl
    import scipy.sparse as sp
    if matrix == "poisson2d":
        A = wl.poisson2d(40)
    elif matrix == "stencil27":
        A = wl.stencil27(9)
    elif matrix == "random_spd":
        A = golden.matrix("cg/random_spd_80/uplo3_pre1")
    else:  # banded, made symmetric positive definite: multi-lane row tiles
        k = int(matrix[6:])
        Bm = wl.banded(1500, k).to_scipy()
        S = (Bm + Bm.T + (4 * k + 8) * sp.identity(1500)).tocsr()
        S.sort_indices()
        A = wl.CsrMatrix(1500, 1500, S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data)
    s = egm.ConjugateGradient(A, loop_mode=loop_mode, chunk_iters=5)
    assert s.multi_rhs_batch() == 4
    s.setTolerance(TOL)
    _check_batched_equals_sequential(egm, s, _columns(wl, A, 6))
    s.close()


def test_batched_with_guess_max_iterations_and_identity_preconditioner(egm):
    from eigen_git_mirror_b200 import workloads as wl
    A = wl.varcoef3d(12)
    B = _columns(wl, A, 5)
    s = egm.ConjugateGradient(A, preconditioner=egm.IdentityPreconditioner)
    s.setTolerance(TOL)
    seq = _check_batched_equals_sequential(egm, s, B)
    X0 = np.asfortranarray(np.stack([x for x, _, _, _ in seq], axis=1))
    X0[:, 0] += 1e-3                      # one column starts off the solution, the others ON it (0 iterations)
    X0[:, 2] = 0.0
    _check_batched_equals_sequential(egm, s, B, X0)
    for k in (0, 1, 3, 10):               # fixed-k trajectories: every live column stops at k with NoConvergence
        s.setMaxIterations(k)
        _check_batched_equals_sequential(egm, s, B)
    s.close()


def test_batched_nonfinite_column_does_not_disturb_the_others(egm):
    from eigen_git_mirror_b200 import workloads as wl
    A = wl.poisson2d(20)
    B = _columns(wl, A, 4)
    B[5, 1] = np.nan
    s = egm.ConjugateGradient(A)
    s.setTolerance(TOL).setMaxIterations(300)
    seq = _check_batched_equals_sequential(egm, s, B)
    assert np.isnan(seq[1][0]).all() and seq[1][1] == 300 and seq[1][3] == egm.NoConvergence
    assert seq[0][3] == egm.Success and np.isfinite(seq[0][0]).all()
    s.close()


def test_handles_without_batched_kernels_solve_column_by_column(egm):
    """Strongly irregular rows (two-phase / long-row tiles) and float factorizations take the sequential path inside
    the same call: same results, same interface."""
    from eigen_git_mirror_b200 import workloads as wl
    import scipy.sparse as sp
    P = wl.powerlaw(3000, 8, seed=5).to_scipy()
    S = (P + P.T + 200.0 * sp.identity(3000)).tocsr()
    S.sort_indices()
    A = wl.CsrMatrix(3000, 3000, S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data)
    s = egm.ConjugateGradient(A)
    st = s.stats()
    assert st["tiles_stream"] + st["tiles_long"] > 0 and s.multi_rhs_batch() == 0
    s.setTolerance(TOL)
    _check_batched_equals_sequential(egm, s, _columns(wl, A, 3))
    s.close()


def test_bicgstab_multi_column_follows_the_reference_loop(egm):
    from eigen_git_mirror_b200 import workloads as wl
    A = wl.convdiff3d(10)
    B = _columns(wl, A, 3)
    s = egm.BiCGSTAB(A)
    s.setTolerance(TOL)
    seq = _sequential(s, B)
    X = s.solve(B)
    for k, (x, it, err, info) in enumerate(seq):
        assert np.array_equal(X[:, k], x)
    assert s.iterations() == seq[-1][1]
    s.close()


def test_batched_256_cubed_matches_reference_fixture_and_pays(egm, fullsize):
    """configs[1] with 4 right-hand sides: column 0 is BASELINE's b, so it must match the unmodified reference's
    solution (tests/golden/fullsize_v1.npz); the batch must cost clearly less than 4 single solves."""
    import torch
    from eigen_git_mirror_b200 import workloads as wl
    from test_gpu_fullsize import assert_close_to_reference, check_against_reference
    A = wl.poisson3d(256)
    S = A.to_scipy()
    cols = [np.asarray(S @ wl.random_vector(A.rows, 12345))] + [np.asarray(S @ wl.random_vector(A.rows, 500 + k)) for k in range(3)]
    B = torch.from_numpy(np.stack(cols, axis=0)).cuda()   # 4 x n, rows contiguous = column-major n x 4
    X = torch.zeros_like(B)
    # WHILE-graph mode for both sides of the bitwise comparison: at this size the persistent kernel AUTO would pick
    # for single columns groups its partial sums differently (see kernels.cuh), the batched path uses the graph kernels
    s = egm.ConjugateGradient(A, loop_mode=1)
    s.setTolerance(TOL)
    s.solve_device_multi(B, X, 4)
    ms_batch = s.stats()["last_solve_ms"]
    x0 = X[0].cpu().numpy()
    assert_close_to_reference(fullsize, "cg_256", "full",
                              [check_against_reference(fullsize, "cg_256", "full", x0, int(s.column_iterations[0]),
                                                       float(s.column_errors[0]), int(s.column_infos[0]))])
    xs = torch.zeros(A.rows, dtype=torch.float64, device="cuda")
    ms_single = 0.0
    for k in range(4):
        s.solve_device(B[k], xs)
        ms_single += s.stats()["last_solve_ms"]
        assert s.iterations() == s.column_iterations[k]
        assert torch.equal(xs, X[k]), k
    print(f"4 columns at 256^3: batched {ms_batch:.1f} ms, sequential {ms_single:.1f} ms, ratio {ms_single / ms_batch:.2f}")
    assert ms_batch < 0.8 * ms_single
    s.close()

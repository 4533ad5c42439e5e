"""GPU: IncompleteLUT / IncompleteCholesky preconditioners through the C ABI, against outputs of the
unmodified reference (tests/golden/golden_v4.npz, make_golden_v4.py).

* z = M^-1 r (level-scheduled triangular solves, csrc/kernels_tri.cuh): BIT-IDENTICAL to IncompleteLUT::solve /
  IncompleteCholesky::solve of the reference -- the factor is the reference's entry for entry (tests/test_factors.py) and
  every row is summed in the reference's order with the reference's roundings;
* ConjugateGradient + IncompleteCholesky, BiCGSTAB + IncompleteLUT, GMRES + IncompleteLUT: info identical, iteration
  count within 2 % (at least +-1; identical on fixed-k trajectories), x within 1e-8 norm-wise, error() <= tol.
The permutation (AMD in the reference) is an input of the product and is taken from the golden file."""
import numpy as np
import pytest

from conftest import _Files, golden_case_names  # noqa: F401

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Golden4:
    def __init__(self):
        self.z = np.load(os.path.join(ROOT, "tests", "golden", "golden_v4.npz"))
        self.keys = list(self.z.keys())

    def cases(self, prefix):
        return [k[: -len("/matrix")] for k in self.keys if k.startswith(prefix + "/") and k.endswith("/matrix")]

    def get(self, case, name, default=None):
        k = f"{case}/{name}"
        if k in self.z:
            v = self.z[k]
            return v.item() if v.ndim == 0 else v
        return default

    def matrix(self, case):
        from eigen_git_mirror_b200.workloads import CsrMatrix
        key = str(self.get(case, "matrix"))
        g = lambda n: self.z[f"{key}/{n}"]
        return CsrMatrix(int(g("rows")), int(g("cols")), g("rowptr"), g("colidx"), g("vals"), 0, key)


G4 = Golden4()


def _perm(case):
    p = G4.get(case, "perm")
    return None if p is None or np.size(p) == 0 else np.ascontiguousarray(p, np.int32)


def _preconditioner(egm, case, kind):
    if kind == "ilut":
        return egm.IncompleteLUT(droptol=float(G4.get(case, "droptol", -1.0)), fillfactor=int(G4.get(case, "fillfactor", 0)),
                                 perm=_perm(case))
    uplo = int(G4.get(case, "uplo", 1))
    return egm.IncompleteCholesky(uplo=(1 if uplo == 3 else uplo), perm=_perm(case) if int(G4.get(case, "ordering", 0)) else None)


@pytest.mark.parametrize("case", G4.cases("precond"))
def test_preconditioner_apply_is_bit_identical(case, egm):
    A = G4.matrix(case)
    kind = str(G4.get(case, "kind"))
    pre = _preconditioner(egm, case, kind)
    s = egm.BiCGSTAB(A, preconditioner=pre)
    assert s.info() == egm.Success == int(G4.get(case, "info"))
    assert s.preconditioner() is pre and int(pre.L.b200s_factors_nnz(pre.handle())) == int(G4.get(case, "factor_nnz"))
    z = s.precondition(G4.get(case, "r"))
    want = G4.get(case, "z")
    assert np.array_equal(z, want), f"{(z != want).sum()} of {z.size} entries differ, max {np.abs(z - want).max():.3e}"
    # the number of kernels is the launch plan's: 2 permute/scale passes + the launches of both stages
    planned = 2 + len(pre.stage(0).launches) + len(pre.stage(1).launches)
    assert s.stats()["last_kernel_launches"] == planned
    # twice the same bits (no dependence on scheduling)
    assert np.array_equal(s.precondition(G4.get(case, "r")), z)
    s.close()


SOLVER_CASES = G4.cases("cg_ichol") + G4.cases("bicgstab_ilut") + G4.cases("gmres_ilut")


@pytest.mark.parametrize("case", SOLVER_CASES)
def test_preconditioned_solver_golden(case, egm):
    A = G4.matrix(case)
    which = str(G4.get(case, "which"))
    if which == "cg_ichol":
        uplo = int(G4.get(case, "uplo"))
        pre = _preconditioner(egm, case, "ichol")
        s = egm.ConjugateGradient(A, uplo=uplo, preconditioner=pre)
    elif which == "bicgstab_ilut":
        s = egm.BiCGSTAB(A, preconditioner=_preconditioner(egm, case, "ilut"))
    else:
        s = egm.GMRES(A, preconditioner=_preconditioner(egm, case, "ilut"))
        s.set_restart(int(G4.get(case, "restart")))
    tol, mi = float(G4.get(case, "tol")), int(G4.get(case, "max_iters"))
    s.setTolerance(tol)
    if mi >= 0:
        s.setMaxIterations(mi)
    b = G4.get(case, "b")
    x = s.solveWithGuess(b, G4.get(case, "x0")) if int(G4.get(case, "has_guess")) else s.solve(b)
    xr, itr = G4.get(case, "x_v3"), int(G4.get(case, "iters_v3"))
    errr, infor = float(G4.get(case, "error_v3")), int(G4.get(case, "info_v3"))
    it4 = int(G4.get(case, "iters_v4"))
    name = case.split("/")[-1]
    nx = np.linalg.norm(xr)
    rel = np.linalg.norm(x - xr) / nx if nx > 0 else np.linalg.norm(x)
    if name == "zero_rhs":
        assert not x.any() and s.iterations() == itr and s.info() == infor and s.error() == errr
    elif name.startswith("traj_k"):
        assert s.iterations() == itr and s.info() == infor, (s.iterations(), itr, s.info(), infor)
        assert rel <= 1e-9, rel
        assert abs(s.error() - errr) <= 1e-6 * errr + 1e-14
    else:
        assert s.info() == infor, (s.info(), infor)
        assert abs(s.iterations() - itr) <= max(1, int(0.02 * itr), 3 * abs(itr - it4)), (s.iterations(), itr)
        assert rel <= 1e-8, rel
        if infor == 0:
            assert s.error() <= s.tolerance()
    s.close()


def test_incomplete_cholesky_cg_at_96_cubed(egm, port):
    """Beyond the goldens: 3-D Poisson 96^3 (885k unknowns, 286 dependency levels), natural ordering.  IC-preconditioned
    CG converges to the known solution in far fewer iterations than Jacobi-preconditioned CG, true residual below tol;
    iterations and x as the CPU restatement of the reference's loop gives them (oracle_cg_precond, pinned bit for bit to
    the unmodified reference by tests/test_factors.py)."""
    from eigen_git_mirror_b200 import workloads as wl
    A = wl.poisson3d(96)
    xt = wl.random_vector(A.rows, 12345)
    b = np.asarray(A.to_scipy() @ xt)
    pre = egm.IncompleteCholesky(uplo=egm.Lower)
    s = egm.ConjugateGradient(A, preconditioner=pre)
    s.setTolerance(1e-10)
    x = s.solve(b)
    assert s.info() == egm.Success and pre.info() == egm.Success
    assert np.linalg.norm(A.to_scipy() @ x - b) <= 2e-10 * np.linalg.norm(b)
    assert np.linalg.norm(x - xt) <= 1e-7 * np.linalg.norm(xt)
    xr, itr, errr, infor = port.cg_factors(A, b, pre, tol=1e-10)
    assert infor == 0 and abs(s.iterations() - itr) <= max(1, 0.02 * itr), (s.iterations(), itr)
    assert np.linalg.norm(x - xr) <= (1e-8 if s.iterations() == itr else 1e-7) * np.linalg.norm(xr)
    j = egm.ConjugateGradient(A)
    j.setTolerance(1e-10)
    j.solve(b)
    assert s.iterations() < 0.6 * j.iterations(), (s.iterations(), j.iterations())
    # multi-column right-hand sides go column by column through the same loop
    B = np.stack([b, 2.0 * b], axis=1)
    X = s.solve(B)
    assert np.array_equal(X[:, 0], x) and s.info() == egm.Success
    # back to Jacobi: set_preconditioner(NULL) restores what factorize() built
    s._hd.check(s._hd.L.b200s_set_preconditioner(s._hd.h, None))
    s._pre_obj = None
    xj = s.solve(b)
    assert s.iterations() == j.iterations() and np.linalg.norm(xj - xt) <= 1e-7 * np.linalg.norm(xt)
    s.close()
    j.close()


def test_ilut_bicgstab_on_convection_diffusion_64_cubed(egm, port):
    from eigen_git_mirror_b200 import workloads as wl
    A = wl.convdiff3d(64)
    xt = wl.random_vector(A.rows, 12345)
    b = np.asarray(A.to_scipy() @ xt)
    pre = egm.IncompleteLUT(droptol=1e-3, fillfactor=4)
    s = egm.BiCGSTAB(A, preconditioner=pre)
    s.setTolerance(1e-10)
    x = s.solve(b)
    assert s.info() == egm.Success
    assert np.linalg.norm(A.to_scipy() @ x - b) <= 2e-10 * np.linalg.norm(b)
    xr, itr, errr, infor = port.bicgstab_factors(A, b, pre, tol=1e-10)   # the reference's loop, restated on the CPU
    assert infor == 0 and abs(s.iterations() - itr) <= 1, (s.iterations(), itr)
    assert np.linalg.norm(x - xr) <= (1e-8 if s.iterations() == itr else 1e-7) * np.linalg.norm(xr)
    j = egm.BiCGSTAB(A)
    j.setTolerance(1e-10)
    j.solve(b)
    assert s.iterations() < j.iterations()
    s.close()
    j.close()


def test_errors(egm):
    from eigen_git_mirror_b200 import workloads as wl
    A = wl.poisson2d(12)
    other = egm.IncompleteLUT(wl.poisson2d(10))
    s = egm.ConjugateGradient(A)
    rc = s._hd.L.b200s_set_preconditioner(s._hd.h, other.handle())
    assert rc == -1 and b"size differs" in s._hd.L.b200s_last_error(s._hd.h)
    Af = A.astype(np.float32)
    f = egm.ConjugateGradient(Af)
    rc = f._hd.L.b200s_set_preconditioner(f._hd.h, egm.IncompleteLUT(A).handle())
    assert rc == -6  # double only
    s.close()
    f.close()


@pytest.mark.parametrize("seed", [42, 20261017])
def test_reference_drivers_on_b200_solvers_with_incomplete_factorizations(seed, egm):
    """oracle/_ref/conformance_precond_b200: test/incomplete_cholesky.cpp, the ILUT lines of test/bicgstab.cpp and
    unsupported/test/gmres.cpp on the b200 classes (include/b200/*.h): the solver's own Eigen::IncompleteCholesky /
    IncompleteLUT object factorizes on the host, its factor is handed to the device (b200s_factors_from_*)."""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "conformance_precond_b200")
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference: make -C oracle conformance)")
    res = subprocess.run([exe, f"s{seed}", "r3"], capture_output=True, text=True, timeout=800)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]


def test_real_matrix_flow_on_a_folder_of_matrix_market_files(tmp_path, egm):
    """The spbenchsolver-style flow (tools/solve_market.py, bench/spbench/spbenchsolver.h:213-300 in the reference): a
    folder with an SPD matrix stored as one triangle and a general matrix with right-hand side and reference solution;
    every solver of the tool's list must converge to the reference solution."""
    import subprocess
    import sys
    import scipy.sparse as sp
    from eigen_git_mirror_b200 import marketio as mio, workloads as wl
    A = wl.poisson3d(14)
    L = sp.tril(A.to_scipy()).tocsr()
    mio.saveMarket(wl.CsrMatrix(A.rows, A.cols, L.indptr.astype(np.int32), L.indices.astype(np.int32), L.data),
                   str(tmp_path / "lap_SPD.mtx"), sym=mio.Symmetric)
    Cm = wl.convdiff3d(12)
    x = wl.random_vector(Cm.rows, 4)
    mio.saveMarket(Cm, str(tmp_path / "cd.mtx"))
    mio.saveMarketVector(np.asarray(Cm.to_scipy() @ x), str(tmp_path / "cd_b.mtx"))
    mio.saveMarketVector(x, str(tmp_path / "cd_x.mtx"))
    out = tmp_path / "out"
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "solve_market.py"), str(tmp_path), "--out", str(out)],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("info=0") == 8, res.stdout  # 5 solvers on the SPD matrix, 3 on the general one
    for which in ("bicgstab", "bicgstab_ilut", "gmres_ilut"):
        got = mio.loadMarketVector(str(out / f"cd_{which}_x.mtx"))
        assert np.linalg.norm(got - x) <= 1e-7 * np.linalg.norm(x), which
    xs = [mio.loadMarketVector(str(out / f"lap_SPD_{w}_x.mtx")) for w in ("cg", "cg_ic", "bicgstab", "bicgstab_ilut", "gmres_ilut")]
    for v in xs[1:]:
        assert np.linalg.norm(v - xs[0]) <= 1e-7 * np.linalg.norm(xs[0])


def test_multicolor_ordering_two_wide_levels(egm, port):
    """An ordering for the GPU (b200s_ordering_multicolor): red-black on the 7-point stencil, so each triangular solve is
    2 grid-wide levels instead of ~3n narrow ones.  Same kernels, same bits as the CPU emulation of the staged solves;
    CG converges to the same solution as with the natural ordering."""
    from eigen_git_mirror_b200 import workloads as wl
    A = wl.poisson3d(40)
    xt = wl.random_vector(A.rows, 12345)
    b = np.asarray(A.to_scipy() @ xt)
    perm, colours = egm.multicolor_ordering(A)
    assert colours == 2
    pre = egm.IncompleteCholesky(uplo=egm.Lower, perm=perm)
    s = egm.ConjugateGradient(A, preconditioner=pre)
    assert [len(pre.stage(w).level_ptr) - 1 for w in (0, 1)] == [2, 2]
    r = wl.random_vector(A.rows, 5)
    z = s.precondition(r)
    assert s.stats()["last_kernel_launches"] == 2 + len(pre.stage(0).launches) + len(pre.stage(1).launches) == 6
    assert np.array_equal(z, port.factors_apply(pre, r))
    s.setTolerance(1e-10)
    x = s.solve(b)
    assert s.info() == egm.Success and np.linalg.norm(x - xt) <= 1e-7 * np.linalg.norm(xt)
    nat = egm.ConjugateGradient(A, preconditioner=egm.IncompleteCholesky(uplo=egm.Lower))
    nat.setTolerance(1e-10)
    nat.solve(b)
    jac = egm.ConjugateGradient(A)
    jac.setTolerance(1e-10)
    jac.solve(b)
    assert nat.iterations() <= s.iterations() < jac.iterations(), (nat.iterations(), s.iterations(), jac.iterations())
    for t in (s, nat, jac):
        t.close()


@pytest.mark.parametrize("seed", [42])
def test_reference_driver_with_the_multicolor_ordering_functor(seed, egm):
    """oracle/_ref/conformance_ordering_b200: check_sparse_spd_solving on b200::ConjugateGradient with
    IncompleteCholesky<double, UpLo, b200::MulticolorOrdering> (include/b200/Ordering.h)."""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "conformance_ordering_b200")
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference: make -C oracle conformance)")
    res = subprocess.run([exe, f"s{seed}", "r3"], capture_output=True, text=True, timeout=800)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]


def test_incomplete_cholesky_cg_at_baseline_size_vs_reference_fixture(egm):
    """BASELINE.json's 3D Poisson 256^3 (16.7 M unknowns) with ConjugateGradient + IncompleteCholesky<double, Lower,
    NaturalOrdering> against the UNMODIFIED reference run on exactly this input (tests/golden/fullsize_v2.npz,
    make_fullsize_v2.py: 239 iterations, 166 s on the CPU): fixed-k trajectories identical in count with x to 1e-9, the
    converged solve within the north-star bars (iterations within 2 %, x within 1e-8), true residual below tol."""
    from eigen_git_mirror_b200 import workloads as wl
    z = np.load(os.path.join(ROOT, "tests", "golden", "fullsize_v2.npz"))
    case = "cg_ichol_256"
    A = wl.poisson3d(256)
    assert A.rows == int(z[f"{case}/rows"]) and A.nnz == int(z[f"{case}/nnz"])
    xt = wl.random_vector(A.rows, 12345)
    b = wl.rhs_from_solution(A, xt)
    idx = z[f"{case}/sample_idx"]
    pre = egm.IncompleteCholesky(uplo=egm.Lower)
    s = egm.ConjugateGradient(A, preconditioner=pre)
    assert s.info() == egm.Success
    s.setTolerance(1e-10)
    for tag in ("k1", "k5", "k20", "full"):
        s.setMaxIterations(int(tag[1:]) if tag != "full" else -1)
        x = s.solve(b)
        itr, errr, infor = int(z[f"{case}/{tag}/iters"]), float(z[f"{case}/{tag}/error"]), int(z[f"{case}/{tag}/info"])
        ref = z[f"{case}/{tag}/x_samples"]
        rel = np.linalg.norm(x[idx] - ref) / np.linalg.norm(ref)
        xn = float(z[f"{case}/{tag}/xnorm"])
        if tag == "full":
            assert s.info() == infor == 0 and abs(s.iterations() - itr) <= max(1, 0.02 * itr), (s.iterations(), itr)
            # same count => same trajectory up to rounding; one iteration more or less moves x by about its own error
            bar = 1e-8 if s.iterations() == itr else 1e-7
            assert s.error() <= 1e-10 and rel <= bar and abs(np.linalg.norm(x) - xn) <= bar * xn, (rel, s.error())
        else:
            assert s.iterations() == itr and s.info() == infor, (tag, s.iterations(), itr)
            assert abs(s.error() - errr) <= 1e-6 * errr and rel <= 1e-9, (tag, s.error(), errr, rel)
    op = egm.SparseOperator(A)
    r = b - op.multiply(x)
    assert np.linalg.norm(r) <= 1.05e-10 * np.linalg.norm(b)
    op.close()
    s.close()

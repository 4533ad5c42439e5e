"""GPU: BASELINE.json's own configurations.

* configs[0] (2D 5-point Poisson 1024^2, the reference's CPU-runnable case) and the 3D problems at 128^3 are solved to
  full convergence by BOTH the CUDA path and the CPU oracle and compared with the north-star tolerances.
* configs[1]/[2] at full size (256^3) are too slow for the oracle inside a test (minutes per solve), so they are checked
  through size-independent properties: SpMV against an independent CSR product and linearity, true residual below tol,
  distance to the known solution, bitwise determinism, and the iteration count the survey measured for the reference
  on this operator (765 for CG at 256^3 with a different random x_true: +-5 %).
"""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1500)]
TOL = 1e-10


def _problem(wl, A, seed=12345):
    x_true = wl.random_vector(A.rows, seed)
    b = np.asarray(A.to_scipy() @ x_true)
    return x_true, b


def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_config0_poisson2d_1024_cg_vs_oracle(egm, port):
    from eigen_git_mirror_b200 import workloads as wl
    A = wl.poisson2d(1024)
    assert A.rows == 1_048_576 and A.nnz == 5_238_784
    x_true, b = _problem(wl, A)
    s = egm.ConjugateGradient(A)
    s.setTolerance(TOL)
    x = s.solve(b)
    xr, itr, errr, infor = port.cg(A, b, tol=TOL)
    assert s.info() == infor == 0
    assert abs(s.iterations() - itr) <= 0.02 * itr, (s.iterations(), itr)
    assert s.error() <= TOL and _rel(x, xr) <= 1e-8
    assert port.true_residual(A, x, b) <= 1.05 * max(TOL, port.true_residual(A, xr, b))


@pytest.mark.parametrize("kind", ["cg", "bicgstab"])
def test_3d_128_vs_oracle(kind, egm, port):
    from eigen_git_mirror_b200 import workloads as wl
    A = wl.poisson3d(128) if kind == "cg" else wl.convdiff3d(128)
    x_true, b = _problem(wl, A)
    s = (egm.ConjugateGradient if kind == "cg" else egm.BiCGSTAB)(A)
    s.setTolerance(TOL)
    x = s.solve(b)
    xr, itr, errr, infor = (port.cg if kind == "cg" else port.bicgstab)(A, b, tol=TOL)
    assert s.info() == infor == 0
    assert abs(s.iterations() - itr) <= max(1, 0.02 * itr), (s.iterations(), itr)
    assert s.error() <= TOL and _rel(x, xr) <= 1e-8, _rel(x, xr)
    assert port.true_residual(A, x, b) <= 2 * TOL


def test_config1_poisson3d_256_properties(egm):
    import scipy.sparse as sp
    from eigen_git_mirror_b200 import workloads as wl
    A = wl.poisson3d(256)
    assert A.rows == 16_777_216 and A.nnz == 117_047_296
    S = A.to_scipy()
    op = egm.SparseOperator(A)
    x1, x2 = wl.random_vector(A.rows, 1), wl.random_vector(A.rows, 2)
    y1, y2 = op.multiply(x1), op.multiply(x2)
    # thread-per-row tiles: the product is bit-identical to a sequential CSR product with separately rounded terms
    # wherever scipy sums in the same order; compare with the scaled tolerance to stay independent of that
    scale = abs(S) @ np.abs(x1)
    assert np.all(np.abs(y1 - S @ x1) <= 1e-13 * scale)
    y12 = op.multiply(x1 + x2)
    assert np.all(np.abs(y12 - (y1 + y2)) <= 4e-13 * (abs(S) @ (np.abs(x1) + np.abs(x2))))
    assert np.array_equal(y1, op.multiply(x1))
    op.close()

    x_true, b = _problem(wl, A)
    s = egm.ConjugateGradient(A)
    s.setTolerance(TOL)
    x = s.solve(b)
    assert s.info() == egm.Success and s.error() <= TOL
    assert abs(s.iterations() - 765) <= 0.05 * 765, s.iterations()
    r = b - S @ x
    assert np.linalg.norm(r) / np.linalg.norm(b) <= 1.05 * TOL
    assert _rel(x, x_true) <= 1e-6
    it, xa = s.iterations(), x
    xb = s.solve(b)
    assert s.iterations() == it and np.array_equal(xa, xb)
    # warm start from the solution: no iteration; one-step restarts as in doc/snippets/BiCGSTAB_step_by_step.cpp
    s.solveWithGuess(b, x)
    assert s.iterations() == 0
    s.setMaxIterations(3)
    x3 = s.solve(b)
    assert s.iterations() == 3 and s.info() == egm.NoConvergence
    s.close()


def test_config2_convdiff3d_256_properties(egm):
    from eigen_git_mirror_b200 import workloads as wl
    A = wl.convdiff3d(256)
    S = A.to_scipy()
    x_true, b = _problem(wl, A)
    s = egm.BiCGSTAB(A)
    s.setTolerance(TOL)
    x = s.solve(b)
    assert s.info() == egm.Success and s.error() <= TOL
    assert abs(s.iterations() - 614) <= 0.15 * 614, s.iterations()  # survey probe: 614 (BiCGSTAB counts are erratic)
    assert np.linalg.norm(b - S @ x) / np.linalg.norm(b) <= 1.05 * TOL
    assert _rel(x, x_true) <= 1e-6
    xb = s.solve(b)
    assert np.array_equal(x, xb)
    s.close()

"""CPU: the plain-C oracle (oracle/oracle.c) is pinned BIT-FOR-BIT to the golden vectors produced by the unmodified
reference (Eigen compiled from /root/reference, tests/golden/make_golden.py), for the ISA variant of this host."""
import numpy as np
import pytest

from conftest import golden_case_names

pytestmark = pytest.mark.timeout(300)


@pytest.mark.parametrize("case", golden_case_names("spmv"))
def test_spmv_bit_exact(case, golden, port, variant):
    A = golden.matrix(case)
    x = golden.get(case, "x")
    y = port.spmv(A, x)  # double: unfused row loop; float: the vectorised-body / fused-epilogue pattern of `variant`
    assert y.dtype == A.vals.dtype
    assert np.array_equal(y, golden.get(case, f"y_{variant}"))


def test_spmv_float_variants_differ_only_in_rounding(golden):
    case = "spmv/banded_500_k16/f32"
    a, b = golden.get(case, "y_v3"), golden.get(case, "y_v4")
    assert np.max(np.abs(a - b)) < 1e-5 and not np.array_equal(a, b)


@pytest.mark.parametrize("case", golden_case_names("symv"))
def test_selfadjoint_product_bit_exact(case, golden, port):
    A = golden.matrix(case)
    y = port.symv(A, golden.get(case, "x"), int(golden.get(case, "uplo")))
    assert np.array_equal(y, golden.get(case, "y"))


def test_jacobi_missing_and_zero_diagonal(golden, port):
    case = "jacobi/missing_diag_40"
    A = golden.matrix(case)
    d = port.jacobi(A)
    assert d[0] == 1.0 and d[5] == 1.0 and d[39] == 1.0 and d[7] == 1.0  # absent / exactly-zero diagonal -> 1
    assert np.array_equal(d * golden.get(case, "r"), golden.get(case, "z"))


def _run(port, golden, case):
    A = golden.matrix(case)
    kind = str(golden.get(case, "kind"))
    x0 = golden.get(case, "x0") if int(golden.get(case, "has_guess")) else None
    kw = dict(x0=x0, tol=float(golden.get(case, "tol")), max_iters=int(golden.get(case, "max_iters")),
              precond=int(golden.get(case, "precond")))
    if kind == "cg":
        return port.cg(A, golden.get(case, "b"), uplo=int(golden.get(case, "uplo")), **kw)
    return port.bicgstab(A, golden.get(case, "b"), **kw)


@pytest.mark.parametrize("case", golden_case_names("cg") + golden_case_names("bicgstab"))
def test_solver_bit_exact(case, golden, port, variant):
    x, it, err, info = _run(port, golden, case)
    assert it == int(golden.get(case, f"iters_{variant}"))
    assert info == int(golden.get(case, f"info_{variant}"))
    gerr = float(golden.get(case, f"error_{variant}"))
    assert err == gerr or (np.isnan(err) and np.isnan(gerr))
    assert np.array_equal(x, golden.get(case, f"x_{variant}"), equal_nan=True)


def test_reference_semantics_visible_in_goldens(golden, variant):
    """The control-flow gotchas of SURVEY.md 8a are present in the reference's own outputs."""
    g = lambda c, n: golden.get(c, f"{n}_{variant}")
    # CG, zero rhs: x = 0, iterations 0, error 0 (ConjugateGradient.h:46-52)
    c = "cg/poisson3d_10/zero_rhs"
    assert g(c, "iters") == 0 and g(c, "error") == 0 and not g(c, "x").any()
    # BiCGSTAB, zero rhs: iterations stay maxIterations (2n), error stays the tolerance (BiCGSTAB.h:47-51)
    c = "bicgstab/convdiff3d_10_g0.5/zero_rhs"
    assert g(c, "iters") == 2 * 1000 and g(c, "error") == 1e-10 and not g(c, "x").any()
    # exact guess: no iteration (ConjugateGradient.h:56-61)
    assert g("cg/poisson3d_10/guess_exact", "iters") == 0
    # maxIterations = k stops after exactly k iterations with NoConvergence
    assert g("cg/poisson3d_10/traj_k5", "iters") == 5 and g("cg/poisson3d_10/traj_k5", "info") == 2


@pytest.mark.parametrize("forced", ["v3", "v4"])
def test_both_isa_variants_are_reproduced(forced, golden, port):
    """Whatever CPU this runs on, the port reproduces BOTH reference builds when told their packet width: 4 doubles /
    float body of 4 for x86-64-v3, 8 doubles / float body of 8 for x86-64-v4."""
    lanes = 4 if forced == "v3" else 8
    for case in ("cg/varcoef3d_10/uplo3_pre1", "cg/random_spd_80/ones", "cg/poisson2d_24/guess"):
        A = golden.matrix(case)
        x0 = golden.get(case, "x0") if int(golden.get(case, "has_guess")) else None
        x, it, err, info = port.cg(A, golden.get(case, "b"), x0=x0, tol=float(golden.get(case, "tol")),
                                   max_iters=int(golden.get(case, "max_iters")), uplo=int(golden.get(case, "uplo")),
                                   precond=int(golden.get(case, "precond")), lanes=lanes)
        assert it == int(golden.get(case, f"iters_{forced}")) and err == float(golden.get(case, f"error_{forced}"))
        assert np.array_equal(x, golden.get(case, f"x_{forced}"))
    # BiCGSTAB: pinned on the AVX-512 build; the AVX2 build contracts the scalar remainder of `x += alpha*y + w*z`
    # differently, which shows on vector lengths that are not a multiple of 4 (random_square_90), not on 512 = 8^3
    for case in (("bicgstab/random_square_90/pre0", "bicgstab/varcoef3d_8/pre1") if forced == "v4"
                 else ("bicgstab/varcoef3d_8/pre1",)):
        A = golden.matrix(case)
        x, it, err, info = port.bicgstab(A, golden.get(case, "b"), tol=float(golden.get(case, "tol")),
                                         max_iters=int(golden.get(case, "max_iters")),
                                         precond=int(golden.get(case, "precond")), lanes=lanes)
        assert it == int(golden.get(case, f"iters_{forced}")) and np.array_equal(x, golden.get(case, f"x_{forced}"))
    case = "spmv/banded_500_k16/f32"
    A = golden.matrix(case)
    y = port.spmv(A, golden.get(case, "x"), blk=4 if forced == "v3" else 8)
    assert np.array_equal(y, golden.get(case, f"y_{forced}"))


# ---------------------------------------------------------------------------------------------- round 2 additions
@pytest.mark.parametrize("case", golden_case_names("bicgstab_restart"))
def test_restart_branch_bit_exact_and_counted(case, golden, port):
    """Systems on which the reference takes BiCGSTAB.h:72-81 (tests/golden/make_golden_v2.py): the port reproduces the
    AVX-512 reference bit for bit at every maxIterations, so its restart count is the reference's."""
    x, it, err, info = port.bicgstab(golden.matrix(case), golden.get(case, "b"), tol=float(golden.get(case, "tol")),
                                     max_iters=int(golden.get(case, "max_iters")),
                                     precond=int(golden.get(case, "precond")), lanes=8)
    assert port.last_restarts == int(golden.get(case, "restarts"))
    assert it == int(golden.get(case, "iters_v4")) and err == float(golden.get(case, "error_v4"))
    assert np.array_equal(x, golden.get(case, "x_v4"))
    # the AVX2 build of the reference takes the same branches
    assert int(golden.get(case, "iters_v3")) == it


def test_restart_goldens_do_restart(golden):
    full = [c for c in golden_case_names("bicgstab_restart") if c.endswith("/full")]
    counts = sorted(int(golden.get(c, "restarts")) for c in full)
    assert len(full) == 7 and counts[0] >= 1 and counts[-1] == 2
    # "reset i only on the first restart" (BiCGSTAB.h:80): the twice-restarting case reports more iterations than
    # its second restart alone would leave
    c = [c for c in full if int(golden.get(c, "restarts")) == 2][0]
    assert int(golden.get(c, "iters_v4")) == 6


@pytest.mark.parametrize("case", golden_case_names("cg_f32") + golden_case_names("bicgstab_f32"))
def test_float_solver_port_tracks_reference(case, golden, port, variant):
    """Float instantiations: the port follows the reference's control flow exactly (iterations, info) on the fixed-k
    and trivial cases and to rounding elsewhere; bit-exactness of float solves is not claimed."""
    x, it, err, info = _run(port, golden, case)
    xr = golden.get(case, f"x_{variant}").astype(np.float64)
    name = case.split("/")[-1]
    if name.startswith("traj_k") or name == "zero_rhs":
        assert it == int(golden.get(case, f"iters_{variant}")) and info == int(golden.get(case, f"info_{variant}"))
    nx = np.linalg.norm(xr)
    if nx > 0:
        assert np.linalg.norm(x.astype(np.float64) - xr) / nx <= 1e-3

"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref, Eigen compiled from
/root/reference) on small seeded problems.  Run in the build container only:

    python tests/golden/make_golden.py

The reference's own tests hold no golden vectors for this path, so parity is pinned by outputs of
the reference itself.  Each case stores the full CSR matrix, right-hand side, optional initial guess, solver
parameters, and the reference's x / iterations() / error() / info() -- both ISA variants of the reference build
(x86-64-v4 = AVX-512 packets, x86-64-v3 = AVX2 packets; they differ in the last bits through Eigen's packet
reductions).  The files are consumed by tests/test_oracle_golden.py (CPU, pins oracle/oracle.c bit-for-bit) and
tests/test_gpu_parity.py (GPU, pins the CUDA path within the north-star tolerances).
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("workloads", os.path.join(ROOT, "eigen-git-mirror_b200", "workloads.py"))
wl = importlib.util.module_from_spec(spec)
sys.modules["workloads"] = wl
spec.loader.exec_module(wl)
from oracle import loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def random_spd(n, density, seed):
    """SPD matrix in the manner of the reference's generate_sparse_spd_problem (test/sparse_solver.h:214-239):
    A = M*M^T with a forced nonzero diagonal in M."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    M = sp.random(n, n, density=density, random_state=rng, data_rvs=lambda k: rng.uniform(-1, 1, k)).tocsr()
    M = M + sp.diags(rng.uniform(0.5, 1.5, n))
    A = (M @ M.T).tocsr()
    A.sort_indices()
    return wl.CsrMatrix(n, n, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64), 0,
                        f"random_spd_{n}")


def random_square(n, density, seed):
    """Nonsymmetric well-conditioned square matrix (test/sparse_solver.h:358-380 analogue, ForceNonZeroDiag)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    M = sp.random(n, n, density=density, random_state=rng, data_rvs=lambda k: rng.uniform(-1, 1, k)).tocsr()
    M = (M + sp.diags(rng.uniform(2.0, 3.0, n) * np.where(rng.random(n) < 0.5, -1, 1))).tocsr()
    M.sort_indices()
    return wl.CsrMatrix(n, n, M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.astype(np.float64), 0,
                        f"random_square_{n}")


def missing_diag(n, seed):
    """SPD-ish matrix where a few diagonal entries are structurally absent / exactly zero: exercises the
    `invdiag = 1` branches of BasicPreconditioners.h:71-74 (SpMV + Jacobi apply only, not a solve)."""
    A = wl.banded(n, 3, seed=seed)
    keep = np.ones(A.nnz, bool)
    rowof = np.repeat(np.arange(n), np.diff(A.rowptr))
    drop_rows = np.array([0, 5, n - 1])
    keep[(rowof == A.colidx) & np.isin(rowof, drop_rows)] = False
    vals = A.vals.copy()
    vals[(rowof == A.colidx) & (rowof == 7)] = 0.0
    counts = np.bincount(rowof[keep], minlength=n)
    rowptr = np.zeros(n + 1, np.int32)
    np.cumsum(counts, out=rowptr[1:])
    return wl.CsrMatrix(n, n, rowptr, A.colidx[keep].copy(), vals[keep].copy(), 0, f"missing_diag_{n}")


def main():
    refs = {"v4": loader.Ref("v4"), "v3": loader.Ref("v3")}
    cases = {}

    def put(name, **kw):
        for k, v in kw.items():
            cases[f"{name}/{k}"] = np.asarray(v)

    def put_matrix(name, A):
        """Matrices are stored once under mat/<A.name>[/f32]; the case records the key."""
        key = f"mat/{A.name}" + ("/f32" if A.vals.dtype == np.float32 else "")
        if f"{key}/rows" not in cases:
            put(key, rows=A.rows, cols=A.cols, rowptr=A.rowptr, colidx=A.colidx, vals=A.vals)
        put(name, matrix=key)

    # ---- SpMV known-answer cases (double and float) -----------------------------------------------------------
    spmv_mats = [wl.poisson2d(20), wl.poisson3d(9), wl.convdiff3d(8), wl.stencil27(7), wl.banded(500, 16),
                 wl.banded(300, 50), wl.powerlaw(700, 8), wl.powerlaw(400, 40, max_row=300), missing_diag(40, 3)]
    for A in spmv_mats:
        for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
            Ad = A.astype(dt)
            x = wl.random_vector(A.cols, 54321, dt)
            name = f"spmv/{A.name}/{tag}"
            put_matrix(name, Ad)
            put(name, x=x, y_v4=refs["v4"].spmv(Ad, x), y_v3=refs["v3"].spmv(Ad, x))
    A = missing_diag(40, 3)
    r = wl.random_vector(40, 99)
    put("jacobi/missing_diag_40", r=r, z=refs["v4"].jacobi_apply(A, r))
    put_matrix("jacobi/missing_diag_40", A)

    # ---- selfadjointView<Lower/Upper> products ------------------------------------------------------------------
    for A in (wl.varcoef3d(6), random_spd(50, 0.08, 5)):
        x = wl.random_vector(A.cols, 4242)
        for uplo in (1, 2):
            name = f"symv/{A.name}/uplo{uplo}"
            put_matrix(name, A)
            put(name, x=x, y=refs["v4"].symv(A, x, uplo), uplo=uplo)

    # ---- solver cases --------------------------------------------------------------------------------------------
    def solver_case(name, kind, A, b, x0=None, tol=1e-10, max_iters=-1, uplo=3, precond=1):
        put_matrix(name, A)
        put(name, b=b, tol=tol, max_iters=max_iters, uplo=uplo, precond=precond, kind=kind,
            has_guess=int(x0 is not None))
        if x0 is not None:
            put(name, x0=x0)
        for v, R in refs.items():
            if kind == "cg":
                x, it, err, info = R.cg(A, b, x0=x0, tol=tol, max_iters=max_iters, uplo=uplo, precond=precond)
            else:
                x, it, err, info = R.bicgstab(A, b, x0=x0, tol=tol, max_iters=max_iters, precond=precond)
            put(name, **{f"x_{v}": x, f"iters_{v}": it, f"error_{v}": err, f"info_{v}": info})

    spd = [wl.poisson2d(24), wl.poisson3d(10), wl.varcoef3d(10), random_spd(80, 0.05, 11)]
    for A in spd:
        xt = wl.random_vector(A.rows, 12345)
        b = wl.rhs_from_solution(A, xt)
        for uplo in (3, 1, 2):
            for pre in (1, 0):
                solver_case(f"cg/{A.name}/uplo{uplo}_pre{pre}", "cg", A, b, uplo=uplo, precond=pre)
        # fixed-k trajectory, default tolerance (epsilon) so that only max_iters stops it
        for k in (0, 1, 2, 5, 10):
            solver_case(f"cg/{A.name}/traj_k{k}", "cg", A, b, tol=-1.0, max_iters=k)
        solver_case(f"cg/{A.name}/ones", "cg", A, np.ones(A.rows))
        solver_case(f"cg/{A.name}/zero_rhs", "cg", A, np.zeros(A.rows))
        solver_case(f"cg/{A.name}/guess", "cg", A, b, x0=xt + 1e-3 * wl.random_vector(A.rows, 77))
        solver_case(f"cg/{A.name}/guess_exact", "cg", A, b, x0=xt)
        solver_case(f"cg/{A.name}/default_tol", "cg", A, b, tol=-1.0, max_iters=-1)
        solver_case(f"cg/{A.name}/loose_tol", "cg", A, b, tol=1e-3)

    nonsym = [wl.convdiff3d(10), wl.convdiff3d(8, gamma=0.9), random_square(90, 0.05, 21), wl.varcoef3d(8)]
    for A in nonsym:
        xt = wl.random_vector(A.rows, 12345)
        b = wl.rhs_from_solution(A, xt)
        for pre in (1, 0):
            solver_case(f"bicgstab/{A.name}/pre{pre}", "bicgstab", A, b, precond=pre)
        for k in (0, 1, 2, 5, 10):
            solver_case(f"bicgstab/{A.name}/traj_k{k}", "bicgstab", A, b, tol=-1.0, max_iters=k)
        solver_case(f"bicgstab/{A.name}/ones", "bicgstab", A, np.ones(A.rows))
        solver_case(f"bicgstab/{A.name}/zero_rhs", "bicgstab", A, np.zeros(A.rows))
        solver_case(f"bicgstab/{A.name}/guess", "bicgstab", A, b, x0=xt + 1e-3 * wl.random_vector(A.rows, 77))
        solver_case(f"bicgstab/{A.name}/guess_exact", "bicgstab", A, b, x0=xt)
        solver_case(f"bicgstab/{A.name}/default_tol", "bicgstab", A, b, tol=-1.0, max_iters=-1)
    # BiCGSTAB restart branch (BiCGSTAB.h:72-81): r0 exactly orthogonal to r after the first step is hard to hit
    # by construction; a guess equal to the solution of a perturbed system gives tiny rho -> restart on some seeds.
    A = wl.convdiff3d(6)
    b = wl.rhs_from_solution(A, wl.random_vector(A.rows, 3))
    solver_case("bicgstab/restart_probe/tiny", "bicgstab", A, b * 1e-160, tol=1e-12)

    path = os.path.join(OUT, "golden_v1.npz")
    np.savez_compressed(path, **cases)
    print(f"wrote {path}: {len(cases)} arrays, {os.path.getsize(path) / 1e6:.2f} MB; reference = "
          f"{refs['v4'].build_info}")


if __name__ == "__main__":
    main()

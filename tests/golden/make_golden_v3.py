"""Generate tests/golden/golden_v3.npz -- outputs of the UNMODIFIED reference (oracle/_ref) for the solvers of SURVEY 8f
rank 3: LeastSquaresConjugateGradient (LeastSquareConjugateGradient.h), MINRES and GMRES (unsupported/Eigen/src/
IterativeSolvers).  Same method and file layout as make_golden.py; build container only:

    python tests/golden/make_golden_v3.py
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


def csr(S, name):
    S = S.tocsr()
    S.sort_indices()
    return wl.CsrMatrix(S.shape[0], S.shape[1], S.indptr.astype(np.int32), S.indices.astype(np.int32),
                        S.data.astype(np.float64), 0, name)


def main():
    import scipy.sparse as sp
    refs = {"v4": loader.Ref("v4"), "v3": loader.Ref("v3")}
    cases = {}

    def put(name, **kw):
        for k, v in kw.items():
            cases[f"{name}/{k}"] = np.asarray(v)

    def put_matrix(name, A):
        key = f"mat/{A.name}"
        if f"{key}/rows" not in cases:
            put(key, rows=A.rows, cols=A.cols, rowptr=A.rowptr, colidx=A.colidx, vals=A.vals)
        put(name, matrix=key)

    def case(name, kind, A, b, x0=None, tol=1e-10, max_iters=-1, **kw):
        put_matrix(name, A)
        put(name, b=b, tol=tol, max_iters=max_iters, kind=kind, has_guess=int(x0 is not None), **kw)
        if x0 is not None:
            put(name, x0=x0)
        for v, R in refs.items():
            x, it, err, info = getattr(R, kind)(A, b, x0=x0, tol=tol, max_iters=max_iters, **kw)
            put(name, **{f"x_{v}": x, f"iters_{v}": it, f"error_{v}": err, f"info_{v}": info})
        return x, it, err, info

    rng = np.random.default_rng(2026)
    # ---- LSCG: rectangular least squares, square nonsymmetric, empty column ------------------------------------
    R1 = sp.random(160, 90, density=0.06, random_state=rng, data_rvs=lambda k: rng.uniform(-1, 1, k)).tolil()
    for j in range(90):
        R1[j, j] = 2.0 + rng.random()
    R2 = sp.random(90, 90, density=0.05, random_state=rng, data_rvs=lambda k: rng.uniform(-1, 1, k)) + sp.diags(rng.uniform(2, 3, 90))
    R3 = R1.copy()
    R3[:, 7] = 0.0  # a structurally empty column: LeastSquareDiagonalPreconditioner leaves invdiag = 0 (row-major)
    for A in (csr(sp.csr_matrix(R1), "ls_rect_160x90"), csr(R2, "ls_square_90"), wl.convdiff3d(7),
              csr(sp.csr_matrix(R3), "ls_rect_emptycol")):
        b = rng.uniform(-1, 1, A.rows)
        for pre in (1, 0):
            x, *_ = case(f"lscg/{A.name}/pre{pre}", "lscg", A, b, precond=pre)
        for k in (0, 1, 2, 5):
            case(f"lscg/{A.name}/traj_k{k}", "lscg", A, b, tol=-1.0, max_iters=k, precond=1)
        case(f"lscg/{A.name}/zero_rhs", "lscg", A, np.zeros(A.rows), precond=1)
        case(f"lscg/{A.name}/guess", "lscg", A, b, x0=x + 1e-3 * rng.standard_normal(A.cols), precond=1)
    # ---- MINRES: SPD, symmetric indefinite, one-triangle views ----------------------------------------------------
    P = wl.poisson2d(14)
    Sind = (P.to_scipy() - 3.3 * sp.identity(P.rows)).tocsr()  # symmetric indefinite (shifted Laplacian)
    for A in (wl.poisson3d(8), wl.varcoef3d(7), csr(Sind, "sym_indefinite_196")):
        b = np.asarray(A.to_scipy() @ wl.random_vector(A.rows, 12345))
        for uplo in (3, 1, 2):
            for pre in ((0, 1) if "indefinite" not in A.name else (0,)):
                x, *_ = case(f"minres/{A.name}/uplo{uplo}_pre{pre}", "minres", A, b, uplo=uplo, precond=pre)
        for k in (0, 1, 2, 5):
            case(f"minres/{A.name}/traj_k{k}", "minres", A, b, tol=-1.0, max_iters=k, uplo=3, precond=0)
        case(f"minres/{A.name}/zero_rhs", "minres", A, np.zeros(A.rows), uplo=3, precond=0)
        case(f"minres/{A.name}/guess", "minres", A, b, x0=x + 1e-3 * rng.standard_normal(A.rows), uplo=3, precond=0)
    # ---- GMRES: nonsymmetric, several restart lengths ---------------------------------------------------------------
    for A in (wl.convdiff3d(8), wl.convdiff3d(6, gamma=0.9), csr(R2, "ls_square_90")):
        b = np.asarray(A.to_scipy() @ wl.random_vector(A.rows, 12345))
        for pre in (1, 0):
            for restart in (30, 5):
                x, *_ = case(f"gmres/{A.name}/pre{pre}_r{restart}", "gmres", A, b, restart=restart, precond=pre)
        for k in (1, 2, 5, 12):
            case(f"gmres/{A.name}/traj_k{k}", "gmres", A, b, tol=-1.0, max_iters=k, restart=5, precond=1)
        case(f"gmres/{A.name}/zero_rhs", "gmres", A, np.zeros(A.rows), restart=30, precond=1)
        case(f"gmres/{A.name}/guess", "gmres", A, b, x0=x + 1e-3 * rng.standard_normal(A.rows), restart=30, precond=1)
    small = csr(sp.csr_matrix(np.array([[4.0, 1, 0, 0], [2, 5, 1, 0], [0, 1, 6, 2], [1, 0, 1, 7]])), "gmres_4x4")
    case("gmres/gmres_4x4/full_krylov", "gmres", small, np.array([1.0, 2, 3, 4]), restart=30, precond=0)  # k == m stop

    path = os.path.join(OUT, "golden_v3.npz")
    np.savez_compressed(path, **cases)
    print(f"wrote {path}: {len(cases)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")
    for k in sorted(cases):
        if k.endswith("iters_v4") and ("/pre" in k or "uplo" in k):
            print(k, int(cases[k]), float(cases[k.replace("iters", "error")]), int(cases[k.replace("iters", "info")]),
                  "v3:", int(cases[k.replace("v4", "v3")]))


if __name__ == "__main__":
    main()

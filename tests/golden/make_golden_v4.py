"""Generate tests/golden/golden_v4.npz -- outputs of the UNMODIFIED reference (oracle/_ref) for SURVEY 8f rank 4:
IncompleteLUT / IncompleteCholesky as preconditioners (IncompleteLUT.h, IncompleteCholesky.h):

  precond/<case>   : z = preconditioner.solve(r) for a fixed r, the permutation the reference chose (AMD), nnz of the factor
  cg_ichol/<case>, bicgstab_ilut/<case>, gmres_ilut/<case> : ConjugateGradient<_, UpLo, IncompleteCholesky<...>>,
                     BiCGSTAB<_, IncompleteLUT>, GMRES<_, IncompleteLUT>: x, iterations(), error(), info()

Generated with the AVX2+FMA build (v3): its IncompleteCholesky::m_scale uses exact square roots, which is what the
product's host factorization reproduces bit for bit (the AVX-512 build approximates them, EIGEN_FAST_MATH); the
AVX-512 build's solver outputs are stored next to them (they agree to rounding).  Build container only:

    python tests/golden/make_golden_v4.py
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
LOWER, UPPER, BOTH = 1, 2, 3


def csr(S, name):
    S = S.tocsr()
    S.sort_indices()
    return wl.CsrMatrix(S.shape[0], S.shape[1], S.indptr.astype(np.int32), S.indices.astype(np.int32),
                        S.data.astype(np.float64), 0, name)


def named(A, name):
    return wl.CsrMatrix(A.rows, A.cols, A.rowptr, A.colidx, A.vals, 0, name)


def main():
    import scipy.sparse as sp
    refs = {"v3": loader.Ref("v3"), "v4": loader.Ref("v4")}
    R = refs["v3"]
    cases = {}

    def put(name, **kw):
        for k, v in kw.items():
            cases[f"{name}/{k}"] = np.asarray(v)

    def put_matrix(name, A):
        key = f"mat/{A.name}"
        if f"{key}/rows" not in cases:
            put(key, rows=A.rows, cols=A.cols, rowptr=A.rowptr, colidx=A.colidx, vals=A.vals)
        put(name, matrix=key)

    rng = np.random.default_rng(2027)
    # a random diagonally dominant nonsymmetric matrix with one very wide dependency level (most rows are independent)
    S = sp.random(3000, 3000, density=0.0006, random_state=rng, data_rvs=lambda k: rng.uniform(-1, 1, k)).tolil()
    S.setdiag(3.0 + rng.random(3000))
    wide = csr(S, "random_wide_3000")
    Ssym = sp.random(700, 700, density=0.01, random_state=rng, data_rvs=lambda k: rng.uniform(-1, 1, k))
    Ssym = (Ssym + Ssym.T).tolil()
    Ssym.setdiag(4.0 + rng.random(700))
    spd = csr(Ssym, "random_spd_700")
    mats = [named(wl.poisson3d(10), "poisson3d_10"), named(wl.convdiff3d(10), "convdiff3d_10"),
            named(wl.poisson2d(40), "poisson2d_40"), named(wl.varcoef3d(9), "varcoef3d_9"), wide, spd]

    # ---- z = M^-1 r ----------------------------------------------------------------------------------------------
    for A in mats:
        r = wl.random_vector(A.rows, 4242)
        for droptol, fill in ((-1.0, 0), (1e-2, 5)):
            name = f"precond/ilut_{A.name}_{'default' if droptol < 0 else 'drop'}"
            rp, ci, va, P, Pinv, info = R.ilut(A, droptol, fill)
            put_matrix(name, A)
            put(name, kind="ilut", droptol=droptol, fillfactor=fill, r=r, z=R.ilut_solve(A, r, droptol, fill), perm=P,
                factor_nnz=len(ci), info=info)
        if A.name in ("convdiff3d_10", "random_wide_3000"):
            continue  # not symmetric
        for uplo in (LOWER, UPPER):
            for ordering in (0, 1):
                name = f"precond/ichol_{A.name}_{'lower' if uplo == LOWER else 'upper'}_{'amd' if ordering else 'natural'}"
                cp, ri, lv, sc, perm, info = R.ichol(A, uplo, ordering)
                put_matrix(name, A)
                put(name, kind="ichol", uplo=uplo, ordering=ordering, r=r, z=R.ichol_solve(A, r, uplo, ordering),
                    perm=perm, factor_nnz=len(ri), info=info)

    # ---- solvers ---------------------------------------------------------------------------------------------------
    def solver_case(name, which, A, b, x0=None, tol=1e-10, max_iters=-1, **kw):
        put_matrix(name, A)
        put(name, b=b, tol=tol, max_iters=max_iters, which=which, has_guess=int(x0 is not None), **kw)
        if x0 is not None:
            put(name, x0=x0)
        if which == "cg_ichol":
            perm = R.ichol(A, kw.get("uplo", LOWER) if kw.get("uplo", LOWER) != BOTH else LOWER, kw.get("ordering", 1))[4]
        else:
            perm = R.ilut(A, kw.get("droptol", -1.0), kw.get("fillfactor", 0))[3]
        put(name, perm=perm)
        for v, Rv in refs.items():
            x, it, err, info = Rv.precond_solver(which, A, b, x0=x0, tol=tol, max_iters=max_iters, **kw)
            put(name, **{f"x_{v}": x, f"iters_{v}": it, f"error_{v}": err, f"info_{v}": info})
            print(f"{name:50s} {v} iters {it:4d} error {err:.3e} info {info}")

    for A in (mats[0], mats[2], mats[3], spd, named(wl.poisson3d(20), "poisson3d_20")):
        xt = wl.random_vector(A.rows, 12345)
        b = np.asarray(A.to_scipy() @ xt)
        for uplo in (LOWER, UPPER, BOTH):
            for ordering in (0, 1):
                tag = {LOWER: "lower", UPPER: "upper", BOTH: "both"}[uplo] + ("_amd" if ordering else "_natural")
                solver_case(f"cg_ichol/{A.name}_{tag}/full", "cg_ichol", A, b, uplo=uplo, ordering=ordering)
        solver_case(f"cg_ichol/{A.name}_lower_amd/traj_k5", "cg_ichol", A, b, max_iters=5, uplo=LOWER, ordering=1)
        solver_case(f"cg_ichol/{A.name}_lower_amd/guess", "cg_ichol", A, b, x0=xt + 1e-3 * wl.random_vector(A.rows, 99),
                    uplo=LOWER, ordering=1)
        solver_case(f"cg_ichol/{A.name}_lower_amd/zero_rhs", "cg_ichol", A, np.zeros(A.rows), uplo=LOWER, ordering=1)
    for A in (mats[1], mats[0], wide, named(wl.convdiff3d(20), "convdiff3d_20")):
        xt = wl.random_vector(A.rows, 12345)
        b = np.asarray(A.to_scipy() @ xt)
        solver_case(f"bicgstab_ilut/{A.name}/full", "bicgstab_ilut", A, b)
        solver_case(f"bicgstab_ilut/{A.name}/drop", "bicgstab_ilut", A, b, droptol=1e-2, fillfactor=5)
        solver_case(f"bicgstab_ilut/{A.name}/traj_k2", "bicgstab_ilut", A, b, max_iters=2, droptol=1e-2, fillfactor=5)
        solver_case(f"bicgstab_ilut/{A.name}/zero_rhs", "bicgstab_ilut", A, np.zeros(A.rows))
        solver_case(f"gmres_ilut/{A.name}/full", "gmres_ilut", A, b, droptol=1e-2, fillfactor=5, restart=30)
        solver_case(f"gmres_ilut/{A.name}/restart4", "gmres_ilut", A, b, droptol=0.1, fillfactor=2, restart=4)

    path = os.path.join(OUT, "golden_v4.npz")
    np.savez_compressed(path, **cases)
    print(f"wrote {path}: {len(cases)} arrays, {os.path.getsize(path) / 1e6:.2f} MB; reference {R.build_info}")


if __name__ == "__main__":
    main()

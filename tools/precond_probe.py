"""Timing probe for the incomplete-factorization preconditioners (one GPU): IncompleteCholesky-CG on 3D Poisson and
IncompleteLUT-BiCGSTAB on convection-diffusion, next to the Jacobi-preconditioned solvers on the same systems.
Prints one JSON object.  No torch import (short GPU calls)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import eigen_git_mirror_b200 as egm  # noqa: E402
from eigen_git_mirror_b200 import workloads as wl  # noqa: E402


def run(kind, n, ordering="natural"):
    A = (wl.poisson3d if kind == "cg" else wl.convdiff3d)(n)
    b = np.asarray(A.to_scipy() @ wl.random_vector(A.rows, 12345))
    t0 = time.perf_counter()
    perm = egm.multicolor_ordering(A)[0] if ordering == "multicolor" else None
    pre = (egm.IncompleteCholesky(uplo=egm.Lower, perm=perm) if kind == "cg"
           else egm.IncompleteLUT(droptol=1e-3, fillfactor=4, perm=perm))
    S = egm.ConjugateGradient if kind == "cg" else egm.BiCGSTAB
    s = S(A, preconditioner=pre)
    setup = time.perf_counter() - t0
    s.setTolerance(1e-10)
    s.solve(b)
    x = s.solve(b)
    st = s.stats()
    res = float(np.linalg.norm(A.to_scipy() @ x - b) / np.linalg.norm(b))
    out = {"solver": S.__name__, "preconditioner": type(pre).__name__, "ordering": ordering, "grid": n, "rows": A.rows, "nnz": A.nnz,
           "iterations": s.iterations(), "error": s.error(), "info": s.info(), "true_residual": res,
           "solve_ms": st["last_solve_ms"], "launches": st["last_kernel_launches"], "setup_s": round(setup, 2),
           "factor_nnz": int(pre.L.b200s_factors_nnz(pre.handle())),
           "levels": [len(pre.stage(w).level_ptr) - 1 for w in (0, 1)],
           "launches_per_apply": 2 + len(pre.stage(0).launches) + len(pre.stage(1).launches)}
    r = wl.random_vector(A.rows, 777)
    s.precondition(r)
    s.precondition(r)
    out["apply_ms"] = s.stats()["last_solve_ms"]
    out["apply_gbs"] = (2 * out["factor_nnz"] * 20 + 4 * A.rows * 8) / (out["apply_ms"] * 1e-3) / 1e9
    s.close()
    j = S(A)
    j.setTolerance(1e-10)
    j.solve(b)
    j.solve(b)
    out["jacobi_iterations"] = j.iterations()
    out["jacobi_solve_ms"] = j.stats()["last_solve_ms"]
    j.close()
    return out


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [128, 64]
    print(json.dumps({"ichol_cg": run("cg", sizes[0]), "ichol_cg_multicolor": run("cg", sizes[0], "multicolor"),
                      "ilut_bicgstab": run("bicg", sizes[1])}))

#!/usr/bin/env python
"""Solve a MatrixMarket system on the GPU -- the real-matrix flow of the reference's spbenchsolver
(/root/reference/bench/spbench/spbenchsolver.h:213-300): symmetric files go to ConjugateGradient with the stored
triangle as UpLo, general files to BiCGSTAB; the right-hand side is <name>_b.mtx if present, else A * ones.

    python tools/solve_market.py matrix.mtx [--tol 1e-10] [--maxit N] [--precond diagonal|identity] [--out x.mtx]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("matrix")
    ap.add_argument("--rhs", default="")
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--maxit", type=int, default=-1)
    ap.add_argument("--precond", default="diagonal", choices=["diagonal", "identity"])
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import eigen_git_mirror_b200 as egm
    from eigen_git_mirror_b200 import marketio as mio

    ok, sym, iscomplex, isvector = mio.getMarketHeader(args.matrix)
    if not ok or iscomplex or isvector:
        raise SystemExit(f"{args.matrix}: need a real coordinate matrix")
    A = mio.loadMarket(args.matrix)
    if A.rows != A.cols:
        raise SystemExit("matrix must be square")
    rhs = args.rhs or args.matrix.replace(".mtx", "_b.mtx")
    rowof = np.repeat(np.arange(A.rows), np.diff(A.rowptr))
    if os.path.exists(rhs):
        b = mio.loadMarketVector(rhs)
    else:
        S = A.to_scipy()
        if sym:  # the file stores one triangle
            import scipy.sparse as sp
            S = S + S.T - sp.diags(S.diagonal())
        b = np.asarray(S @ np.ones(A.rows))
    pre = egm.DiagonalPreconditioner if args.precond == "diagonal" else egm.IdentityPreconditioner
    t0 = time.perf_counter()
    if sym:
        uplo = egm.Lower if np.all(A.colidx <= rowof) else egm.Upper if np.all(A.colidx >= rowof) else (egm.Lower | egm.Upper)
        solver = egm.ConjugateGradient(A, uplo=uplo, preconditioner=pre)
        name = f"ConjugateGradient<UpLo={uplo}>"
    else:
        solver = egm.BiCGSTAB(A, preconditioner=pre)
        name = "BiCGSTAB"
    t_setup = time.perf_counter() - t0
    solver.setTolerance(args.tol)
    if args.maxit >= 0:
        solver.setMaxIterations(args.maxit)
    x = solver.solve(b)
    st = solver.stats()
    print(f"{name} n={A.rows} nnz={A.nnz} iterations={solver.iterations()} error={solver.error():.3e} "
          f"info={solver.info()} setup={t_setup:.3f}s solve={st['last_solve_ms']:.3f}ms")
    if args.out:
        mio.saveMarketVector(x, args.out)


if __name__ == "__main__":
    main()

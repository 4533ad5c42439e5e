#!/usr/bin/env python
"""Solve MatrixMarket systems on the GPU -- the real-matrix flow of the reference's spbenchsolver
(/root/reference/bench/spbench/spbenchsolver.h:213-300, .cpp): for every matrix of a folder (or one file) run the
iterative solvers the reference runs -- SPD matrices: ConjugateGradient with Jacobi and with IncompleteCholesky;
all matrices: BiCGSTAB with Jacobi and with IncompleteLUT, GMRES with IncompleteLUT -- and report iterations, time and
the error against the reference solution when there is one (name_x.mtx, or the random solution behind a generated
right-hand side).  Right-hand side: name_b.mtx, else A * random (MatrixMarketIterator.h:112-133).

    python tools/solve_market.py <folder | matrix.mtx> [--tol 1e-10] [--maxit N] [--solvers cg,cg_ic,bicgstab,bicgstab_ilut,gmres_ilut]
                                 [--ordering natural|rcm] [--droptol 1e-3] [--fillfactor 10] [--out DIR]

The incomplete factorizations take their fill-reducing permutation as an input; the reference uses AMD, which is
outside this path: `--ordering rcm` uses scipy's reverse Cuthill-McKee, `natural` none.
"""
import argparse
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ordering_perm(A, kind):
    """Permutation in the reference's convention (m_P / m_perm .indices(): row i of A becomes row perm[i])."""
    if kind == "natural":
        return None
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    S = A.to_scipy()
    order = reverse_cuthill_mckee((abs(S) + abs(S.T)).tocsr(), symmetric_mode=True)  # new position k holds old row order[k]
    perm = np.empty(A.rows, np.int32)
    perm[order] = np.arange(A.rows, dtype=np.int32)
    return perm


def run_one(egm, name, A, b, refx, sym, args):
    rows = []
    perm = ordering_perm(A, args.ordering)
    spd = sym != 0
    for which in args.solvers:
        if which in ("cg", "cg_ic") and not spd:
            continue
        t0 = time.perf_counter()
        try:
            if which == "cg":
                s = egm.ConjugateGradient(A)
            elif which == "cg_ic":
                s = egm.ConjugateGradient(A, preconditioner=egm.IncompleteCholesky(uplo=egm.Lower, perm=perm))
            elif which == "bicgstab":
                s = egm.BiCGSTAB(A)
            elif which == "bicgstab_ilut":
                s = egm.BiCGSTAB(A, preconditioner=egm.IncompleteLUT(droptol=args.droptol, fillfactor=args.fillfactor, perm=perm))
            elif which == "gmres_ilut":
                s = egm.GMRES(A, preconditioner=egm.IncompleteLUT(droptol=args.droptol, fillfactor=args.fillfactor, perm=perm))
            else:
                raise SystemExit(f"unknown solver {which}")
        except egm.B200Error as e:
            rows.append((which, "setup failed: " + str(e)))
            continue
        setup = time.perf_counter() - t0
        if s.info() != egm.Success:  # spbenchsolver.h: "The preconditioner / factorization failed"
            rows.append((which, f"compute() reported info={s.info()} (factorization failed)"))
            s.close()
            continue
        s.setTolerance(args.tol)
        if args.maxit >= 0:
            s.setMaxIterations(args.maxit)
        x = s.solve(b)
        ms = s.stats()["last_solve_ms"]
        res = float(np.linalg.norm(A.to_scipy() @ x - b) / max(np.linalg.norm(b), 1e-300))
        err = float(np.linalg.norm(x - refx) / np.linalg.norm(refx)) if refx.size == x.size and np.linalg.norm(refx) > 0 else None
        rows.append((which, f"iterations={s.iterations():6d} error()={s.error():.3e} info={s.info()} setup={setup:7.3f}s "
                            f"solve={ms:10.3f}ms true_residual={res:.3e}" + (f" rel_err_vs_refX={err:.3e}" if err is not None else "")))
        if args.out:
            from eigen_git_mirror_b200 import marketio as mio
            mio.saveMarketVector(x, os.path.join(args.out, f"{name}_{which}_x.mtx"))
        s.close()
    print(f"== {name}: n={A.rows} nnz={A.nnz} {'SPD/symmetric' if spd else 'general'}")
    for which, line in rows:
        print(f"   {which:14s} {line}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path", help="a folder of MatrixMarket files, or one matrix file")
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--maxit", type=int, default=-1)
    ap.add_argument("--solvers", default="cg,cg_ic,bicgstab,bicgstab_ilut,gmres_ilut")
    ap.add_argument("--ordering", default="rcm", choices=["natural", "rcm"])
    ap.add_argument("--droptol", type=float, default=1e-3)
    ap.add_argument("--fillfactor", type=int, default=10)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    args.solvers = [s for s in args.solvers.split(",") if s]
    import eigen_git_mirror_b200 as egm
    from eigen_git_mirror_b200 import marketio as mio

    folder, tmp = args.path, None
    if os.path.isfile(args.path):  # one file: present it to the iterator as a folder of its own
        tmp = tempfile.mkdtemp()
        base = os.path.basename(args.path)[:-4]
        for suffix in ("", "_b", "_x"):
            src = os.path.join(os.path.dirname(os.path.abspath(args.path)), base + suffix + ".mtx")
            if os.path.exists(src):
                shutil.copy(src, tmp)
        folder = tmp
    if args.out:
        os.makedirs(args.out, exist_ok=True)
    it = mio.MatrixMarketIterator(folder)
    if not it.isFolderValid():
        raise SystemExit(f"{folder}: not a folder")
    n = 0
    while it:
        A = it.matrix()
        if A.rows == A.cols:
            b = it.rhs()
            run_one(egm, it.matname(), A, b, it.refX(), it.sym(), args)
            n += 1
        it.next()
    if tmp:
        shutil.rmtree(tmp, ignore_errors=True)
    if n == 0:
        raise SystemExit("no real square MatrixMarket matrix found")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""CPU study behind DESIGN 7e: how the permutation given to IncompleteCholesky / IncompleteLUT shapes what the GPU has to
run -- dependency levels and kernel launches per apply -- against what it costs in iterations.  Orderings: natural, the
reference's AMD (taken from oracle/_ref when it is available), reverse Cuthill-McKee (scipy), multi-colour
(b200s_ordering_multicolor).  Iteration counts come from the CPU restatement of the reference's loops (oracle/, pinned bit
for bit to the reference).  Prints JSON lines; no GPU is used.

    python tools/ordering_study.py [n_poisson=64] [n_convdiff=48] > profiles/r2_ordering_levels.jsonl
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import eigen_git_mirror_b200 as egm  # noqa: E402
from eigen_git_mirror_b200 import workloads as wl  # noqa: E402
from oracle import loader  # noqa: E402
from solve_market import ordering_perm  # noqa: E402


def orderings(A, kind, R):
    out = {"natural": None, "rcm": ordering_perm(A, "rcm"), "multicolor": egm.multicolor_ordering(A)[0]}
    if R is not None:
        out["amd (reference)"] = R.ichol(A, 1, 1)[4] if kind == "ichol" else R.ilut(A)[3]
    return out


def main():
    n_p = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n_c = int(sys.argv[2]) if len(sys.argv) > 2 else 48
    port = loader.port()
    R = loader.Ref("v3") if loader.ref_available() else None
    for kind, A, name in (("ichol", wl.poisson3d(n_p), f"3D Poisson {n_p}^3, CG + IncompleteCholesky<Lower>"),
                          ("ilut", wl.convdiff3d(n_c), f"3D convection-diffusion {n_c}^3, BiCGSTAB + IncompleteLUT(1e-3, 10)")):
        b = np.asarray(A.to_scipy() @ wl.random_vector(A.rows, 12345))
        jac = (port.cg if kind == "ichol" else port.bicgstab)(A, b, tol=1e-10)[1]
        for oname, perm in orderings(A, kind, R).items():
            t0 = time.time()
            pre = (egm.IncompleteCholesky(A, uplo=egm.Lower, perm=perm) if kind == "ichol"
                   else egm.IncompleteLUT(A, droptol=1e-3, fillfactor=10, perm=perm))
            t_fact = time.time() - t0
            st = [pre.stage(w) for w in (0, 1)]
            solve = port.cg_factors if kind == "ichol" else port.bicgstab_factors
            _, it, err, info = solve(A, b, pre, tol=1e-10)
            widths = [np.diff(s.level_ptr) for s in st]
            print(json.dumps({
                "problem": name, "rows": A.rows, "ordering": oname, "factor_nnz": int(pre.L.b200s_factors_nnz(pre.handle())),
                "levels": [len(s.level_ptr) - 1 for s in st], "widest_level": [int(w.max()) for w in widths],
                "median_level": [int(np.median(w)) for w in widths],
                "launches_per_apply": 2 + sum(len(s.launches) for s in st),
                "iterations": int(it), "info": int(info), "jacobi_iterations": int(jac),
                "host_factorization_s": round(t_fact, 3)}), flush=True)


if __name__ == "__main__":
    main()

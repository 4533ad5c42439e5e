#!/usr/bin/env python
"""Runs a few CG (or BiCGSTAB) iterations at full size through the C ABI: the target of the ncu captures of the fused
vector kernels.  --loop-mode 3 launches every kernel as an ordinary stream launch (easiest to profile)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--solver", default="cg")
ap.add_argument("--matrix", default="")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--loop-mode", type=int, default=3)
args = ap.parse_args()
import eigen_git_mirror_b200 as egm
from eigen_git_mirror_b200 import workloads as wl
A = wl.poisson2d(args.n) if args.matrix == "poisson2d" else (wl.poisson3d if args.solver == "cg" else wl.convdiff3d)(args.n)
b = np.asarray(A.to_scipy() @ wl.random_vector(A.rows, 12345))
S = egm.ConjugateGradient if args.solver == "cg" else egm.BiCGSTAB
s = S(A, loop_mode=args.loop_mode)
s.setTolerance(1e-10).setMaxIterations(args.iters)
x = s.solve(b)
for _ in range(2):
    x = s.solve(b)
st = s.stats()
print(args.solver, "n", args.n, "mode", args.loop_mode, "iters", s.iterations(), "error", s.error(), "launches",
      st["last_kernel_launches"], "solve_ms %.3f" % st["last_solve_ms"],
      "us/iter %.2f" % (1e3 * st["last_solve_ms"] / max(1, s.iterations())), "l2_persist", st["l2_persist"],
      "evict_first", st["evict_first"], s.timeline())

#!/usr/bin/env python
"""Multi-column CG probe: K right-hand sides at n^3 through b200s_cg_solve_multi_device_f64, against K single solves.

    python tools/multi_probe.py --n 256 --cols 4 [--iters 0] [--loop-mode 1]
Prints one JSON line: device ms of the batch and of the sequential solves, column-iterations per second of both."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--cols", type=int, default=4)
ap.add_argument("--iters", type=int, default=0)
ap.add_argument("--loop-mode", type=int, default=1)
ap.add_argument("--skip-single", action="store_true")
args = ap.parse_args()
import torch
import eigen_git_mirror_b200 as egm
from eigen_git_mirror_b200 import workloads as wl

A = wl.poisson3d(args.n)
S = A.to_scipy()
B = torch.from_numpy(np.stack([np.asarray(S @ wl.random_vector(A.rows, 12345 + 7 * k)) for k in range(args.cols)])).cuda()
X = torch.zeros_like(B)
s = egm.ConjugateGradient(A, loop_mode=args.loop_mode)
s.setTolerance(1e-10)
if args.iters:
    s.setMaxIterations(args.iters)
s.solve_device_multi(B, X, args.cols)
s.solve_device_multi(B, X, args.cols)
ms_b = s.stats()["last_solve_ms"]
col_iters = int(np.sum(s.column_iterations))
out = {"n": args.n, "cols": args.cols, "batch": s.multi_rhs_batch(), "batched_ms": ms_b, "column_iterations": col_iters,
       "batched_col_it_per_s": col_iters / (ms_b * 1e-3), "iterations": s.column_iterations.tolist(),
       "launches": s.stats()["last_kernel_launches"]}
if not args.skip_single:
    xs = torch.zeros(A.rows, dtype=torch.float64, device="cuda")
    ms_s = 0.0
    for k in range(args.cols):
        s.solve_device(B[k], xs)
        ms_s += s.stats()["last_solve_ms"]
        assert torch.equal(xs, X[k]) or args.loop_mode == 4
    out.update(sequential_ms=ms_s, sequential_col_it_per_s=col_iters / (ms_s * 1e-3), speedup=ms_s / ms_b)
print(json.dumps(out))

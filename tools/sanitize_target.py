#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer's synccheck / racecheck: every SpMV tile flavour (1..32 lanes per row,
two-phase "stream" tiles, rows longer than a tile) in double and float, CG through plain stream launches and through
the persistent cooperative kernel (hand-rolled grid barriers), BiCGSTAB through stream launches.  The WHILE-graph mode
is left to memcheck: the tools do not support device-side cudaGraphSetConditional."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eigen_git_mirror_b200 as egm  # noqa: E402
from eigen_git_mirror_b200 import workloads as wl  # noqa: E402

for A in (wl.poisson3d(10), wl.banded(600, 16), wl.banded(400, 100), wl.stencil27(8), wl.powerlaw(900, 12),
          wl.powerlaw(300, 60, max_row=3000)):
    for dt in (np.float64, np.float32):
        op = egm.SparseOperator(A.astype(dt))
        x = wl.random_vector(A.cols, 1, dt)
        y = op.multiply(x)
        ref = A.astype(dt).to_scipy() @ x
        assert np.allclose(y, ref, rtol=1e-4 if dt == np.float32 else 1e-11, atol=1e-4 if dt == np.float32 else 1e-11)
        st = op.stats()
        print(A.name, np.dtype(dt).name, "lanes", st["tiles_by_lanes"], "stream", st["tiles_stream"], "long", st["tiles_long"])
        op.close()
A = wl.varcoef3d(12)
b = wl.rhs_from_solution(A, wl.random_vector(A.rows, 12345))
for mode in (egm.solvers.LOOP_STREAM, egm.solvers.LOOP_PERSISTENT):
    for dt in (np.float64, np.float32):
        s = egm.ConjugateGradient(A.astype(dt), loop_mode=mode)
        s.setTolerance(1e-10 if dt == np.float64 else 1e-5)
        x = s.solve(b.astype(dt))
        assert s.info() == 0, (mode, dt, s.info(), s.error())
        print("cg mode", mode, np.dtype(dt).name, "iters", s.iterations())
        s.close()
C = wl.convdiff3d(10)
bc = wl.rhs_from_solution(C, wl.random_vector(C.rows, 12345))
s = egm.BiCGSTAB(C, loop_mode=egm.solvers.LOOP_STREAM)
s.setTolerance(1e-10)
s.solve(bc)
assert s.info() == 0
print("bicgstab iters", s.iterations())
print("sanitize target ok")

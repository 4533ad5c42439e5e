"""eigen-git-mirror_b200 -- a B200-native (sm_100a) drop-in for ONE path of Eigen: CSR SpMV and the
ConjugateGradient / BiCGSTAB loops with Jacobi or identity preconditioning.

Layout
  csrc/         hand-written CUDA kernels + the C ABI (include/b200sparse.h) -> lib/libb200sparse.so
  solvers.py    host-side mirror of Eigen's solver interface over that C ABI (ctypes)
  planning.py   GPU-free view of the partition / halo / tile plan (host logic, testable on CPU)
  workloads.py  synthetic CSR matrices of BASELINE.json
  preconditioners.py  IncompleteLUT / IncompleteCholesky: host factorization, GPU application (csrc/kernels_tri.cuh)
  marketio.py   MatrixMarket I/O with the reference's semantics (SparseExtra/MarketIO.h), host-side
  build.py      nvcc build (in-tree)

Importing the package does not load the CUDA library; the first solver / operator does, and raises if it is missing.
"""
from . import workloads  # noqa: F401
from . import solvers  # noqa: F401
from .solvers import (GMRES, MINRES, BiCGSTAB, Communicator, ConjugateGradient, DiagonalPreconditioner,  # noqa: F401
                      IdentityPreconditioner, InvalidInput, LeastSquaresConjugateGradient, Lower, NoConvergence,
                      NumericalIssue, SparseOperator, Success, Upper, device_count, partition_rows)
from .preconditioners import IncompleteCholesky, IncompleteLUT, multicolor_ordering  # noqa: F401
from ._lib import B200Error  # noqa: F401

__all__ = ["ConjugateGradient", "BiCGSTAB", "LeastSquaresConjugateGradient", "MINRES", "GMRES", "SparseOperator", "Communicator", "partition_rows", "device_count",
           "Lower", "Upper", "Success", "NumericalIssue", "NoConvergence", "InvalidInput", "DiagonalPreconditioner",
           "IdentityPreconditioner", "IncompleteLUT", "IncompleteCholesky", "multicolor_ordering", "B200Error", "workloads"]

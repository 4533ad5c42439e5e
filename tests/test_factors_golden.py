"""CPU: the committed golden vectors of the incomplete-factorization preconditioners (tests/golden/golden_v4.npz, outputs
of the unmodified reference) against the product's host factorization + the CPU emulation of the device's staged apply
(oracle/oracle.c).  Needs neither a GPU nor oracle/_ref: this is the same plumbing tests/test_gpu_precond.py uses, with
the emulation in place of the kernels."""
import numpy as np
import pytest

from test_gpu_precond import G4, _preconditioner


class _Egm:  # the two classes _preconditioner needs, without touching the device
    from eigen_git_mirror_b200.preconditioners import IncompleteCholesky, IncompleteLUT  # noqa: F401


@pytest.mark.parametrize("case", G4.cases("precond"))
def test_golden_preconditioner_apply_emulated(case, port):
    A = G4.matrix(case)
    pre = _preconditioner(_Egm, case, str(G4.get(case, "kind")))
    pre.compute(A)
    assert pre.info() == int(G4.get(case, "info")) == 0
    assert int(pre.L.b200s_factors_nnz(pre.handle())) == int(G4.get(case, "factor_nnz"))
    z = port.factors_apply(pre, G4.get(case, "r"), order="level")
    assert np.array_equal(z, G4.get(case, "z"))


def test_golden_file_has_the_expected_cases():
    assert len(G4.cases("precond")) == 28 and len(G4.cases("cg_ichol")) == 45
    assert len(G4.cases("bicgstab_ilut")) == 16 and len(G4.cases("gmres_ilut")) == 8
    wide = [c for c in G4.cases("precond") if "random_wide" in c]
    assert wide, "a case with a grid-wide dependency level must exist"

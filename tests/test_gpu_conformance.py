"""GPU: API conformance through the C++ binding.  oracle/_ref/conformance_b200 holds the reference's own test
drivers (test/sparse_solver.h: check_sparse_spd_solving / check_sparse_square_solving -- dense, sparse and
multi-column right-hand sides, solveWithGuess, analyzePattern+factorize, Map / uncompressed / expression inputs,
matrix constructor, results against dense Householder QR at 1e-6) instantiated on b200::ConjugateGradient and
b200::BiCGSTAB (include/b200/IterativeSolvers.h).  They are compiled in the build container against the reference
headers (make -C oracle conformance) and travel prebuilt; here they only run."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


@pytest.mark.parametrize("seed", [42, 20261017])
def test_reference_driver_on_b200_solvers(seed, egm):
    exe = os.path.join(ROOT, "oracle", "_ref", "conformance_b200")
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference: make -C oracle conformance)")
    res = subprocess.run([exe, f"s{seed}", "r3"], capture_output=True, text=True, timeout=800)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.parametrize("world", [2, 4])
def test_row_partitioned_path_from_cpp(world, egm):
    """oracle/_ref/distributed_b200: one forked process per GPU, b200::ConjugateGradient / BiCGSTAB / SparseOperator with
    setDistributed() and a shared-memory all-gather in place of MPI, against the reference's CPU solvers."""
    exe = os.path.join(ROOT, "oracle", "_ref", "distributed_b200")
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference: make -C oracle distributed)")
    if egm.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    res = subprocess.run([exe, str(world), "24"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and f"all {world} ranks ok" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]

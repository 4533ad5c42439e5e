"""CPU: include/b200/Ordering.h (b200::MulticolorOrdering) inside the reference's own CPU classes --
oracle/ordering_check.cpp: Eigen::ConjugateGradient + Eigen::IncompleteCholesky<double, Lower, b200::MulticolorOrdering>
converges to the known solution, and the factor analysed by the GPU-free ABI has 2 dependency levels per triangular solve
on the 7-point stencil (natural ordering: 3n - 2).  Built where /root/reference exists (make -C oracle ordering), shipped
prebuilt otherwise; no device is touched."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multicolor_ordering_functor_in_eigen_classes():
    exe = os.path.join(ROOT, "oracle", "_ref", "ordering_check")
    if os.path.isdir("/root/reference/Eigen"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ordering"], check=True, capture_output=True)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference)")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "ordering_check ok" in res.stdout, res.stdout + res.stderr
    assert "levels 2 2; natural ordering levels 34 34" in res.stdout

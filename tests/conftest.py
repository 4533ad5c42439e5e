import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) device; run with `-m gpu` on the GPU box")


class _Files:
    """Several .npz files behind one mapping (keys are disjoint by construction)."""

    def __init__(self, paths):
        self.files = [np.load(p) for p in paths]
        self.where = {k: f for f in self.files for k in f.keys()}

    def keys(self):
        return self.where.keys()

    def __contains__(self, k):
        return k in self.where

    def __getitem__(self, k):
        return self.where[k][k]


class Golden:
    """tests/golden/golden_v{1,2,3}.npz: outputs of the unmodified reference (see tests/golden/make_golden.py,
    make_golden_v2.py, make_golden_v3.py)."""

    def __init__(self):
        self.z = _Files([os.path.join(ROOT, "tests", "golden", f)
                         for f in ("golden_v1.npz", "golden_v2.npz", "golden_v3.npz")])
        self.keys = list(self.z.keys())

    def cases(self, prefix):
        seen = []
        for k in self.keys:
            if k.startswith(prefix + "/") and k.endswith("/matrix"):
                seen.append(k[: -len("/matrix")])
        return seen

    def get(self, case, name, default=None):
        k = f"{case}/{name}"
        if k in self.z:
            v = self.z[k]
            return v.item() if v.ndim == 0 else v
        return default

    def matrix(self, case):
        from eigen_git_mirror_b200.workloads import CsrMatrix
        key = str(self.get(case, "matrix"))
        g = lambda n: self.z[f"{key}/{n}"]
        return CsrMatrix(int(g("rows")), int(g("cols")), g("rowptr"), g("colidx"), g("vals"), 0, key)


class Fullsize:
    """tests/golden/fullsize_v1.npz: the unmodified reference on BASELINE.json's own configurations at full size
    (tests/golden/make_fullsize.py): iterations, error(), ||x|| and 4,096 sampled entries of x, for the converged
    solve and for the fixed-k trajectories."""

    def __init__(self):
        self.z = np.load(os.path.join(ROOT, "tests", "golden", "fullsize_v1.npz"))

    def get(self, case, name):
        v = self.z[f"{case}/{name}"]
        return v.item() if v.ndim == 0 else v

    def has(self, case, name):
        return f"{case}/{name}" in self.z


@pytest.fixture(scope="session")
def fullsize():
    return Fullsize()


_golden = None


@pytest.fixture(scope="session")
def golden():
    global _golden
    if _golden is None:
        _golden = Golden()
    return _golden


def golden_case_names(prefix):
    global _golden
    if _golden is None:
        _golden = Golden()
    return _golden.cases(prefix)


@pytest.fixture(scope="session")
def port():
    from oracle import loader
    return loader.port()


@pytest.fixture(scope="session")
def variant():
    """Which ISA variant of the reference build this host reproduces (oracle/loader.host_lanes)."""
    from oracle import loader
    return "v4" if loader.host_lanes() == 8 else "v3"


@pytest.fixture(scope="session")
def egm():
    """The product package, on a machine with a B200.  GPU tests must not silently fall back to anything."""
    import eigen_git_mirror_b200 as m
    if m.device_count() < 1:
        pytest.skip("no sm_100 device visible (GPU tests run under gpurun)")
    return m

"""CPU: the C-ABI library loads, exports every symbol include/b200sparse.h declares, and refuses to run without a
device (no CPU fallback).  No compute call is made."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200sparse.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200s_[a-z0-9_]+)\s*\(", text)) - {"b200s_allgather_fn"})


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for s in ("b200s_create", "b200s_destroy", "b200s_analyze_pattern", "b200s_factorize_f64", "b200s_spmv_f64",
              "b200s_cg_solve_f64", "b200s_bicgstab_solve_f64", "b200s_cg_solve_f32", "b200s_bicgstab_solve_f32",
              "b200s_cg_solve_device_f32", "b200s_get_stats", "b200s_last_error", "b200s_plan_probe_span"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from eigen_git_mirror_b200 import _lib
    L = _lib.lib()
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing
    assert L.b200s_version() == 200


def test_no_torch_types_in_signatures():
    text = open(os.path.join(ROOT, "include", "b200sparse.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # comments may mention the Python host
    assert "torch" not in code and "at::" not in code and "#include <cuda" not in code


def test_create_fails_loudly_without_a_device():
    from eigen_git_mirror_b200 import _lib
    L = _lib.lib()
    if L.b200s_device_count() > 0:
        pytest.skip("a B200 is present")
    h = C.c_void_p()
    rc = L.b200s_create(None, C.byref(h))
    assert rc == -2 and not h.value  # B200S_ERR_NO_DEVICE
    assert b"no CPU path" in L.b200s_last_error(None)
    import eigen_git_mirror_b200 as egm
    with pytest.raises(egm.B200Error):
        egm.ConjugateGradient()


def test_product_does_not_reference_the_oracle():
    """The product package must never import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "eigen-git-mirror_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f in ("kernels.cuh",), f
    # kernels.cuh mentions oracle/oracle_body.h in a comment only
    k = open(os.path.join(pkg, "csrc", "kernels.cuh")).read()
    for line in k.splitlines():
        if "oracle" in line.lower():
            assert line.strip().startswith("//")

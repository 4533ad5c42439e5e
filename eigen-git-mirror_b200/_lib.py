"""ctypes binding of include/b200sparse.h.  Loading fails loudly: there is no Python / CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# B200S_LIB lets an experiment load an alternatively built library (e.g. another CTA size); default is the in-tree build
LIB_PATH = os.environ.get("B200S_LIB") or os.path.join(HERE, "lib", "libb200sparse.so")

ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
        ("spmv_impl", C.c_int32), ("loop_mode", C.c_int32), ("chunk_iters", C.c_int32), ("tile_nnz", C.c_int32),
        ("tile_rows", C.c_int32), ("reserved0", C.c_int32), ("allgather", ALLGATHER_FN), ("allgather_ctx", C.c_void_p),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("world", C.c_int32), ("rank", C.c_int32), ("device", C.c_int32),
        ("rows", C.c_int64), ("cols", C.c_int64), ("nnz", C.c_int64), ("ghosts", C.c_int64), ("halo_send", C.c_int64),
        ("tiles", C.c_int32), ("tiles_boundary", C.c_int32), ("tiles_by_lanes", C.c_int32 * 6),
        ("tiles_stream", C.c_int32), ("tiles_long", C.c_int32),
        ("spmv_grid", C.c_int32), ("spmv_block", C.c_int32), ("spmv_smem_bytes", C.c_int32), ("spmv_stages", C.c_int32),
        ("vec_grid", C.c_int32), ("vec_block", C.c_int32), ("loop_mode", C.c_int32), ("evict_first", C.c_int32),
        ("sm_count", C.c_int32),
        ("last_solve_ms", C.c_double), ("last_h2d_ms", C.c_double), ("last_d2h_ms", C.c_double),
        ("last_kernel_launches", C.c_int64), ("last_iterations", C.c_int64), ("last_spmv_count", C.c_int64),
        ("device_bytes", C.c_int64), ("last_restarts", C.c_int64), ("last_nonfinite", C.c_int32),
        ("last_comm_error", C.c_int32), ("l2_persist", C.c_int32), ("reserved1", C.c_int32),
    ]

    def as_dict(self):
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out


_lib = None


def lib() -> C.CDLL:
    """The loaded libb200sparse.so.  Raises if it has not been built (python -m eigen_git_mirror_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first (python __graft_entry__.py or "
            "python -m eigen_git_mirror_b200.build).  This package has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    H = C.c_void_p
    L.b200s_version.restype = C.c_int
    L.b200s_device_count.restype = C.c_int
    L.b200s_create.argtypes = [C.POINTER(Config), C.POINTER(H)]
    L.b200s_destroy.argtypes = [H]
    L.b200s_destroy.restype = None
    L.b200s_last_error.argtypes = [H]
    L.b200s_last_error.restype = C.c_char_p
    L.b200s_analyze_pattern.argtypes = [H, i64, i64, i64, vp, vp, vp, C.c_int, vp]
    L.b200s_factorize_f64.argtypes = [H, vp, C.c_int]
    L.b200s_factorize_f32.argtypes = [H, vp, C.c_int]
    for sfx in ("f64", "f32"):
        getattr(L, f"b200s_spmv_{sfx}").argtypes = [H, vp, vp]
        getattr(L, f"b200s_spmv_device_{sfx}").argtypes = [H, vp, vp, C.c_int, C.POINTER(C.c_float)]
    solve_args = [H, vp, vp, C.c_int, dbl, i64, C.POINTER(i64), C.POINTER(dbl), C.POINTER(C.c_int)]
    for sfx in ("f64", "f32"):
        for name in ("b200s_cg_solve_", "b200s_bicgstab_solve_", "b200s_cg_solve_device_",
                     "b200s_bicgstab_solve_device_"):
            getattr(L, name + sfx).argtypes = solve_args
    multi_args = [H, i64, vp, i64, vp, i64, C.c_int, dbl, i64, vp, vp, vp]
    L.b200s_cg_solve_multi_f64.argtypes = multi_args
    L.b200s_cg_solve_multi_device_f64.argtypes = multi_args
    L.b200s_multi_rhs_batch.argtypes = [H]
    out3 = [C.POINTER(i64), C.POINTER(dbl), C.POINTER(C.c_int)]
    L.b200s_lscg_solve_f64.argtypes = [H, H, vp, vp, C.c_int, dbl, i64, C.c_int, C.c_int] + out3
    L.b200s_minres_solve_f64.argtypes = [H, vp, vp, C.c_int, dbl, i64] + out3
    L.b200s_gmres_solve_f64.argtypes = [H, vp, vp, C.c_int, dbl, i64, i64] + out3
    F = C.c_void_p
    csr = [i64, vp, vp, vp]
    L.b200s_ilut_f64.argtypes = csr + [dbl, C.c_int, vp, C.POINTER(F)]
    L.b200s_ichol_f64.argtypes = csr + [C.c_int, dbl, vp, C.POINTER(F)]
    L.b200s_factors_from_ilut_f64.argtypes = csr + [vp, C.POINTER(F)]
    L.b200s_factors_from_ichol_f64.argtypes = csr + [vp, vp, C.POINTER(F)]
    L.b200s_ordering_multicolor.argtypes = [i64, vp, vp, vp]
    L.b200s_factors_destroy.argtypes = [F]
    L.b200s_factors_destroy.restype = None
    for name in ("info", "kind"):
        getattr(L, f"b200s_factors_{name}").argtypes = [F]
    for name in ("size", "nnz", "perm_size"):
        getattr(L, f"b200s_factors_{name}").argtypes = [F]
        getattr(L, f"b200s_factors_{name}").restype = i64
    L.b200s_factors_get.argtypes = [F, vp, vp, vp, vp, vp]
    L.b200s_factors_stage_sizes.argtypes = [F, C.c_int, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32),
                                            C.POINTER(i32)]
    L.b200s_factors_stage.argtypes = [F, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    L.b200s_factors_permscale.argtypes = [F, vp, vp, vp, vp, vp]
    L.b200s_set_preconditioner.argtypes = [H, F]
    L.b200s_precond_apply_f64.argtypes = [H, vp, vp]
    L.b200s_get_stats.argtypes = [H, C.POINTER(Stats)]
    L.b200s_get_invdiag_f64.argtypes = [H, vp]
    L.b200s_get_timeline.argtypes = [H, vp, C.c_int]
    L.b200s_get_residual_history.argtypes = [H, vp, i64]
    L.b200s_get_residual_history.restype = i64
    L.b200s_plan_probe.argtypes = [C.POINTER(Config), i64, i64, i64, vp, vp, vp, vp, vp, i64, vp, i64, vp, vp,
                                   C.POINTER(Stats)]
    L.b200s_plan_probe.restype = i64
    L.b200s_plan_probe_csr.argtypes = [i64, i64, vp, vp, vp, C.c_int, vp, vp, vp, i64]
    L.b200s_plan_probe_csr.restype = i64
    L.b200s_plan_probe_selfadjoint.argtypes = [C.POINTER(Config), i64, i64, i64, vp, vp, vp, C.c_int, vp, vp, vp, vp,
                                               vp, i64]
    L.b200s_plan_probe_selfadjoint.restype = i64
    L.b200s_plan_probe_span.argtypes = [i64, i64, vp, vp, vp, C.c_int]
    L.b200s_plan_probe_span.restype = i64
    _lib = L
    return L


class B200Error(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"b200sparse status {status}: {message}")
        self.status = status

"""oracle/loader.py -- TEST INFRASTRUCTURE: ctypes bindings of the two CPU checkers.

* ``port()``  -> oracle/_build/liboracle.so : the plain-C restatement (oracle.c / oracle_body.h).
* ``ref()``   -> oracle/_ref/libeigen_ref_v{4,3}.so : the unmodified reference (Eigen) compiled from
  /root/reference by oracle/Makefile in the build container; travels prebuilt to the GPU box.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this module.  The product package
(eigen-git-mirror_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LOWER, UPPER, BOTH = 1, 2, 3
IDENTITY, JACOBI = 0, 1

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def _cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def host_lanes(dtype=np.float64) -> int:
    """Packet width (in scalars) that the loaded oracle/_ref variant reduces with."""
    v4 = {"avx512f", "avx512dq", "avx512vl", "avx512bw", "avx512cd"} <= _cpu_flags()
    doubles = 8 if v4 else 4
    return doubles * (2 if np.dtype(dtype) == np.float32 else 1)


def build_port():
    subprocess.run(["make", "-C", HERE, "port"], check=True, capture_output=True)


def build_ref(reference="/root/reference"):
    """Build oracle/_ref from the reference sources where they lie; no-op when they are absent (GPU box)."""
    if os.path.isdir(os.path.join(reference, "Eigen")):
        subprocess.run(["make", "-C", HERE, "ref", f"REF={reference}"], check=True, capture_output=True)
        return True
    return False


class _OracleFactors(C.Structure):
    """oracle.c: oracle_factors."""
    _fields_ = [("n", C.c_int64), ("pre_gather", C.c_void_p), ("pre_scale", C.c_void_p), ("post_gather", C.c_void_p),
                ("post_scale", C.c_void_p), ("rowptr", C.c_void_p * 2), ("colidx", C.c_void_p * 2),
                ("vals", C.c_void_p * 2), ("diag", C.c_void_p * 2), ("order", C.c_void_p * 2), ("fused", C.c_int32 * 2),
                ("work", C.c_void_p)]


class Port:
    def __init__(self):
        path = os.path.join(HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build_port()
        self.lib = L = C.CDLL(path)
        for sfx, fp, real in (("f64", _f64p, C.c_double), ("f32", _f32p, C.c_float)):
            getattr(L, f"oracle_spmv_{sfx}").argtypes = [C.c_int64, _i32p, _i32p, fp, fp, fp]
            getattr(L, f"oracle_spmv_{sfx}").restype = None
            getattr(L, f"oracle_spmv_blk_{sfx}").argtypes = [C.c_int64, _i32p, _i32p, fp, fp, fp, C.c_int]
            getattr(L, f"oracle_spmv_blk_{sfx}").restype = None
            getattr(L, f"oracle_symv_{sfx}").argtypes = [C.c_int64, _i32p, _i32p, fp, fp, fp, C.c_int]
            getattr(L, f"oracle_symv_{sfx}").restype = None
            getattr(L, f"oracle_jacobi_factorize_{sfx}").argtypes = [C.c_int64, _i32p, _i32p, fp, fp]
            getattr(L, f"oracle_jacobi_factorize_{sfx}").restype = None
            getattr(L, f"oracle_dot_{sfx}").argtypes = [fp, fp, C.c_int64, C.c_int]
            getattr(L, f"oracle_dot_{sfx}").restype = real
            getattr(L, f"oracle_cg_{sfx}").argtypes = [C.c_int64, _i32p, _i32p, fp, fp, fp, real, C.c_int64, C.c_int,
                                                      C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(real),
                                                      C.POINTER(C.c_int)]
            getattr(L, f"oracle_cg_{sfx}").restype = None
            getattr(L, f"oracle_bicgstab_{sfx}").argtypes = [C.c_int64, _i32p, _i32p, fp, fp, fp, real, C.c_int64,
                                                            C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(real),
                                                            C.POINTER(C.c_int)]
            getattr(L, f"oracle_bicgstab_{sfx}").restype = None
        L.oracle_true_residual_f64.argtypes = [C.c_int64, _i32p, _i32p, _f64p, _f64p, _f64p]
        L.oracle_true_residual_f64.restype = C.c_double
        L.oracle_tri_stage_f64.argtypes = [C.c_int64, _i32p, _i32p, _f64p, C.c_void_p, _i32p, C.c_int, _f64p]
        L.oracle_tri_stage_f64.restype = None
        L.oracle_permute_scale_f64.argtypes = [C.c_int64, _f64p, C.c_void_p, C.c_void_p, _f64p]
        L.oracle_permute_scale_f64.restype = None
        L.oracle_cg_factors_f64.argtypes = [C.c_int64, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_double, C.c_int64, C.c_int,
                                            C.c_int, C.POINTER(_OracleFactors), C.POINTER(C.c_int64),
                                            C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.oracle_cg_factors_f64.restype = None
        L.oracle_bicgstab_factors_f64.argtypes = [C.c_int64, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_double, C.c_int64,
                                                  C.c_int, C.POINTER(_OracleFactors), C.POINTER(C.c_int64),
                                                  C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.oracle_bicgstab_factors_f64.restype = None

    @staticmethod
    def _sfx(a):
        return "f32" if a.dtype == np.float32 else "f64"

    def spmv(self, A, x, blk=None):
        """blk: see oracle_spmv_blk; default = the pattern of the oracle/_ref variant this host loads
        (0 for double; for float 8 on AVX-512 hosts, 4 otherwise)."""
        if blk is None:
            blk = 0 if A.vals.dtype == np.float64 else host_lanes(np.float32) // 2
        y = np.empty(A.rows, dtype=A.vals.dtype)
        getattr(self.lib, f"oracle_spmv_blk_{self._sfx(A.vals)}")(A.rows, A.rowptr, A.colidx, A.vals,
                                                                 np.ascontiguousarray(x, A.vals.dtype), y, blk)
        return y

    def symv(self, A, x, uplo):
        y = np.empty(A.rows, dtype=A.vals.dtype)
        getattr(self.lib, f"oracle_symv_{self._sfx(A.vals)}")(A.rows, A.rowptr, A.colidx, A.vals,
                                                             np.ascontiguousarray(x, A.vals.dtype), y, uplo)
        return y

    def jacobi(self, A):
        d = np.empty(A.rows, dtype=A.vals.dtype)
        getattr(self.lib, f"oracle_jacobi_factorize_{self._sfx(A.vals)}")(A.rows, A.rowptr, A.colidx, A.vals, d)
        return d

    def dot(self, a, b, lanes):
        return getattr(self.lib, f"oracle_dot_{self._sfx(a)}")(a, b, a.shape[0], lanes)

    def _solve(self, fn, A, b, x0, tol, max_iters, extra):
        dt = A.vals.dtype
        real = C.c_float if dt == np.float32 else C.c_double
        x = np.zeros(A.rows, dt) if x0 is None else np.array(x0, dtype=dt, copy=True)
        it, err, info = C.c_int64(0), real(0), C.c_int(0)
        fn(A.rows, A.rowptr, A.colidx, A.vals, np.ascontiguousarray(b, dt), x, tol, max_iters, *extra,
           C.byref(it), C.byref(err), C.byref(info))
        return x, it.value, err.value, info.value

    def cg(self, A, b, x0=None, tol=-1.0, max_iters=-1, uplo=BOTH, precond=JACOBI, lanes=None):
        lanes = host_lanes(A.vals.dtype) if lanes is None else lanes
        return self._solve(getattr(self.lib, f"oracle_cg_{self._sfx(A.vals)}"), A, b, x0, tol, max_iters,
                           (uplo, precond, lanes))

    def bicgstab(self, A, b, x0=None, tol=-1.0, max_iters=-1, precond=JACOBI, lanes=None):
        lanes = host_lanes(A.vals.dtype) if lanes is None else lanes
        return self._solve(getattr(self.lib, f"oracle_bicgstab_{self._sfx(A.vals)}"), A, b, x0, tol, max_iters,
                           (precond, lanes))

    @property
    def last_restarts(self) -> int:
        """Restarts (BiCGSTAB.h:72-81) taken by the last bicgstab() call of this port."""
        return int(C.c_int64.in_dll(self.lib, "oracle_last_restarts").value)

    def factors_apply(self, pre, r, order="level", fused=None):
        """z = M^-1 r with the stages of an IncompleteLUT / IncompleteCholesky object of the product
        (eigen_git_mirror_b200.preconditioners), on the CPU.  order "level": rows in the device's level order;
        "natural": the reference's sequential substitution order.  fused: None = as each stage says, or a bool /
        pair of bools to override (probing what the reference's compiled loops do)."""
        n = pre.rows()
        pg, ps, qg, qs = pre.permscale()
        vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        keep = [np.ascontiguousarray(a) if a is not None else None for a in (pg, ps, qg, qs)]
        x = np.empty(n)
        self.lib.oracle_permute_scale_f64(n, np.ascontiguousarray(r, np.float64), vp(keep[0]), vp(keep[1]), x)
        for which in (0, 1):
            st = pre.stage(which)
            if order == "level":
                rows = np.ascontiguousarray(st.level_rows, np.int32)
            else:
                rows = np.arange(n, dtype=np.int32) if which == 0 else np.arange(n - 1, -1, -1, dtype=np.int32)
            dg = None if st.diag is None else np.ascontiguousarray(st.diag)
            fu = st.fused if fused is None else (fused[which] if isinstance(fused, (tuple, list)) else fused)
            self.lib.oracle_tri_stage_f64(n, st.rowptr, np.ascontiguousarray(st.colidx), np.ascontiguousarray(st.vals),
                                          vp(dg), rows, int(fu), x)
        z = np.empty(n)
        self.lib.oracle_permute_scale_f64(n, x, vp(keep[2]), vp(keep[3]), z)
        return z

    @staticmethod
    def _factors_struct(pre):
        """oracle_factors for an IncompleteLUT / IncompleteCholesky object of the product (natural row order, i.e. the
        reference's sequential substitutions); returns (struct, keep-alive list)."""
        n = pre.rows()
        keep = []

        def ptr(a, dt):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dt)
            keep.append(a)
            return a.ctypes.data

        f = _OracleFactors()
        f.n = n
        pg, ps, qg, qs = pre.permscale()
        f.pre_gather, f.pre_scale = ptr(pg, np.int32), ptr(ps, np.float64)
        f.post_gather, f.post_scale = ptr(qg, np.int32), ptr(qs, np.float64)
        for w in (0, 1):
            st = pre.stage(w)
            f.rowptr[w], f.colidx[w], f.vals[w] = ptr(st.rowptr, np.int32), ptr(st.colidx, np.int32), ptr(st.vals, np.float64)
            f.diag[w] = ptr(st.diag, np.float64)
            order = np.arange(n, dtype=np.int32) if w == 0 else np.arange(n - 1, -1, -1, dtype=np.int32)
            f.order[w] = ptr(order, np.int32)
            f.fused[w] = int(st.fused)
        f.work = ptr(np.zeros(max(n, 1)), np.float64)
        return f, keep

    def cg_factors(self, A, b, pre, x0=None, tol=-1.0, max_iters=-1, uplo=BOTH, lanes=None):
        """ConjugateGradient<_, uplo, IncompleteCholesky> (or any factor preconditioner `pre` of the product)."""
        lanes = host_lanes(np.float64) if lanes is None else lanes
        f, keep = self._factors_struct(pre)
        x = np.zeros(A.rows) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
        it, err, info = C.c_int64(0), C.c_double(0), C.c_int(0)
        self.lib.oracle_cg_factors_f64(A.rows, A.rowptr, A.colidx, np.ascontiguousarray(A.vals, np.float64),
                                       np.ascontiguousarray(b, np.float64), x, tol, max_iters, uplo, lanes, C.byref(f),
                                       C.byref(it), C.byref(err), C.byref(info))
        return x, it.value, err.value, info.value

    def bicgstab_factors(self, A, b, pre, x0=None, tol=-1.0, max_iters=-1, lanes=None):
        """BiCGSTAB<_, IncompleteLUT> (or any factor preconditioner `pre` of the product)."""
        lanes = host_lanes(np.float64) if lanes is None else lanes
        f, keep = self._factors_struct(pre)
        x = np.zeros(A.rows) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
        it, err, info = C.c_int64(0), C.c_double(0), C.c_int(0)
        self.lib.oracle_bicgstab_factors_f64(A.rows, A.rowptr, A.colidx, np.ascontiguousarray(A.vals, np.float64),
                                             np.ascontiguousarray(b, np.float64), x, tol, max_iters, lanes, C.byref(f),
                                             C.byref(it), C.byref(err), C.byref(info))
        return x, it.value, err.value, info.value

    def true_residual(self, A, x, b):
        return self.lib.oracle_true_residual_f64(A.rows, A.rowptr, A.colidx, A.vals.astype(np.float64),
                                                 np.ascontiguousarray(x, np.float64),
                                                 np.ascontiguousarray(b, np.float64))


class Ref:
    """The unmodified reference (Eigen) behind oracle/ref_eigen.cpp."""

    def __init__(self, variant=None):
        if variant is None:
            variant = "v4" if host_lanes() == 8 else "v3"
        self.variant = variant
        path = os.path.join(HERE, "_ref", f"libeigen_ref_{variant}.so")
        if not os.path.exists(path):
            if not build_ref() or not os.path.exists(path):
                raise FileNotFoundError(f"{path}: build it with `make -C oracle ref` where /root/reference exists")
        self.lib = L = C.CDLL(path)
        L.eigref_max_threads.restype = C.c_int
        L.eigref_build_info.restype = C.c_char_p
        dp, ip64, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int)
        for sfx, fp in (("f64", _f64p), ("f32", _f32p)):
            getattr(L, f"eigref_spmv_{sfx}").argtypes = [C.c_int64, C.c_int64, C.c_int64, _i32p, _i32p, fp, fp, fp,
                                                        C.c_int, C.c_int, dp]
            getattr(L, f"eigref_cg_{sfx}").argtypes = [C.c_int64, C.c_int64, _i32p, _i32p, fp, fp, fp, C.c_int,
                                                      C.c_double, C.c_int64, C.c_int, C.c_int, C.c_int, ip64, dp, ip,
                                                      dp, dp]
            getattr(L, f"eigref_bicgstab_{sfx}").argtypes = [C.c_int64, C.c_int64, _i32p, _i32p, fp, fp, fp, C.c_int,
                                                            C.c_double, C.c_int64, C.c_int, C.c_int, ip64, dp, ip, dp,
                                                            dp]
        if hasattr(L, "eigref_lscg_f64"):
            L.eigref_lscg_f64.argtypes = [C.c_int64, C.c_int64, C.c_int64, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_int,
                                          C.c_double, C.c_int64, C.c_int, ip64, dp, ip]
            L.eigref_minres_f64.argtypes = [C.c_int64, C.c_int64, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_int, C.c_double,
                                            C.c_int64, C.c_int, C.c_int, ip64, dp, ip]
            L.eigref_gmres_f64.argtypes = [C.c_int64, C.c_int64, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_int, C.c_double,
                                           C.c_int64, C.c_int64, C.c_int, ip64, dp, ip]
        if hasattr(L, "eigref_ilut_f64"):
            csr = [C.c_int64, C.c_int64, _i32p, _i32p, _f64p]
            L.eigref_ilut_f64.argtypes = csr + [C.c_double, C.c_int, _i32p, _i32p, _f64p, C.c_int64, _i32p, _i32p, ip]
            L.eigref_ilut_f64.restype = C.c_int64
            L.eigref_ilut_solve_f64.argtypes = csr + [C.c_double, C.c_int, _f64p, _f64p]
            L.eigref_ichol_f64.argtypes = csr + [C.c_int, C.c_int, C.c_double, _i32p, _i32p, _f64p, C.c_int64, _f64p,
                                                 _i32p, ip, ip]
            L.eigref_ichol_f64.restype = C.c_int64
            L.eigref_ichol_solve_f64.argtypes = csr + [C.c_int, C.c_int, C.c_double, _f64p, _f64p]
            L.eigref_precond_solver_f64.argtypes = [C.c_int] + csr + [_f64p, _f64p, C.c_int, C.c_double, C.c_int64,
                                                                      C.c_int, C.c_int, C.c_double, C.c_int, C.c_int64,
                                                                      ip64, dp, ip]
        L.eigref_symv_f64.argtypes = [C.c_int64, C.c_int64, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_int]
        L.eigref_jacobi_f64.argtypes = [C.c_int64, C.c_int64, _i32p, _i32p, _f64p, _f64p, _f64p]

    @property
    def max_threads(self):
        return self.lib.eigref_max_threads()

    @property
    def build_info(self):
        return self.lib.eigref_build_info().decode()

    @staticmethod
    def _sfx(a):
        return "f32" if a.dtype == np.float32 else "f64"

    def spmv(self, A, x, threads=1, reps=1):
        y = np.empty(A.rows, dtype=A.vals.dtype)
        best = C.c_double(0)
        rc = getattr(self.lib, f"eigref_spmv_{self._sfx(A.vals)}")(A.rows, A.cols, A.nnz, A.rowptr, A.colidx, A.vals,
                                                                  np.ascontiguousarray(x, A.vals.dtype), y, threads,
                                                                  reps, C.byref(best))
        assert rc == 0
        self.last_seconds = best.value
        return y

    def symv(self, A, x, uplo):
        y = np.zeros(A.rows, dtype=np.float64)
        assert self.lib.eigref_symv_f64(A.rows, A.nnz, A.rowptr, A.colidx, A.vals, x, y, uplo) == 0
        return y

    def jacobi_apply(self, A, r):
        z = np.empty(A.rows, dtype=np.float64)
        assert self.lib.eigref_jacobi_f64(A.rows, A.nnz, A.rowptr, A.colidx, A.vals, r, z) == 0
        return z

    def _solve(self, fn, A, b, x0, tol, max_iters, extra, threads):
        dt = A.vals.dtype
        x = np.zeros(A.rows, dt) if x0 is None else np.array(x0, dtype=dt, copy=True)
        it, err, info = C.c_int64(0), C.c_double(0), C.c_int(0)
        ts, tv = C.c_double(0), C.c_double(0)
        rc = fn(A.rows, A.nnz, A.rowptr, A.colidx, A.vals, np.ascontiguousarray(b, dt), x, int(x0 is not None),
                tol, max_iters, *extra, threads, C.byref(it), C.byref(err), C.byref(info), C.byref(ts), C.byref(tv))
        assert rc == 0
        self.last_setup_seconds, self.last_solve_seconds = ts.value, tv.value
        return x, it.value, err.value, info.value

    def cg(self, A, b, x0=None, tol=-1.0, max_iters=-1, uplo=BOTH, precond=JACOBI, threads=1):
        return self._solve(getattr(self.lib, f"eigref_cg_{self._sfx(A.vals)}"), A, b, x0, tol, max_iters,
                           (uplo, precond), threads)

    def bicgstab(self, A, b, x0=None, tol=-1.0, max_iters=-1, precond=JACOBI, threads=1):
        return self._solve(getattr(self.lib, f"eigref_bicgstab_{self._sfx(A.vals)}"), A, b, x0, tol, max_iters,
                           (precond,), threads)

    def _krylov(self, fn, A, b, x0, ncols_x, args):
        x = np.zeros(ncols_x) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
        it, err, info = C.c_int64(0), C.c_double(0), C.c_int(0)
        rc = fn(*args(np.ascontiguousarray(b, np.float64), x), C.byref(it), C.byref(err), C.byref(info))
        assert rc == 0
        return x, it.value, err.value, info.value

    def lscg(self, A, b, x0=None, tol=-1.0, max_iters=-1, precond=JACOBI):
        return self._krylov(self.lib.eigref_lscg_f64, A, b, x0, A.cols,
                            lambda bb, x: (A.rows, A.cols, A.nnz, A.rowptr, A.colidx, A.vals, bb, x, int(x0 is not None),
                                           tol, max_iters, precond))

    def minres(self, A, b, x0=None, tol=-1.0, max_iters=-1, uplo=LOWER, precond=IDENTITY):
        return self._krylov(self.lib.eigref_minres_f64, A, b, x0, A.rows,
                            lambda bb, x: (A.rows, A.nnz, A.rowptr, A.colidx, A.vals, bb, x, int(x0 is not None), tol,
                                           max_iters, uplo, precond))

    def gmres(self, A, b, x0=None, tol=-1.0, max_iters=-1, restart=30, precond=JACOBI):
        return self._krylov(self.lib.eigref_gmres_f64, A, b, x0, A.rows,
                            lambda bb, x: (A.rows, A.nnz, A.rowptr, A.colidx, A.vals, bb, x, int(x0 is not None), tol,
                                           max_iters, restart, precond))


    # ---- IncompleteLUT / IncompleteCholesky of the unmodified reference ----
    def ilut(self, A, droptol=-1.0, fillfactor=0):
        """(lu_rowptr, lu_colidx, lu_vals, P, Pinv, info): m_lu, m_P, m_Pinv of IncompleteLUT<double>(A)."""
        n = A.rows
        rp, P, Pinv, info = np.zeros(n + 1, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32), C.c_int(0)
        cap = 1
        while True:
            ci, va = np.zeros(cap, np.int32), np.zeros(cap, np.float64)
            nz = self.lib.eigref_ilut_f64(n, A.nnz, A.rowptr, A.colidx, A.vals, droptol, fillfactor, rp, ci, va, cap, P,
                                          Pinv, C.byref(info))
            if nz <= cap:
                return rp, ci[:nz].copy(), va[:nz].copy(), P, Pinv, info.value
            cap = int(nz)

    def ilut_solve(self, A, r, droptol=-1.0, fillfactor=0):
        z = np.zeros(A.rows)
        self.lib.eigref_ilut_solve_f64(A.rows, A.nnz, A.rowptr, A.colidx, A.vals, droptol, fillfactor,
                                       np.ascontiguousarray(r, np.float64), z)
        return z

    def ichol(self, A, uplo=LOWER, ordering=0, shift=-1.0):
        """(L_colptr, L_rowidx, L_vals, scale, perm, info): m_L, m_scale, m_perm (empty = natural ordering)."""
        n = A.rows
        cp, scale, perm = np.zeros(n + 1, np.int32), np.zeros(n), np.zeros(max(n, 1), np.int32)
        psz, info = C.c_int(0), C.c_int(0)
        cap = 1
        while True:
            ri, va = np.zeros(cap, np.int32), np.zeros(cap, np.float64)
            nz = self.lib.eigref_ichol_f64(n, A.nnz, A.rowptr, A.colidx, A.vals, uplo, ordering, shift, cp, ri, va, cap,
                                           scale, perm, C.byref(psz), C.byref(info))
            assert nz >= 0
            if nz <= cap:
                return cp, ri[:nz].copy(), va[:nz].copy(), scale, perm[:psz.value].copy(), info.value
            cap = int(nz)

    def ichol_solve(self, A, r, uplo=LOWER, ordering=0, shift=-1.0):
        z = np.zeros(A.rows)
        self.lib.eigref_ichol_solve_f64(A.rows, A.nnz, A.rowptr, A.colidx, A.vals, uplo, ordering, shift,
                                        np.ascontiguousarray(r, np.float64), z)
        return z

    def precond_solver(self, which, A, b, x0=None, tol=-1.0, max_iters=-1, uplo=LOWER, ordering=1, droptol=-1.0,
                       fillfactor=0, restart=0):
        """which: "cg_ichol", "bicgstab_ilut" or "gmres_ilut" -- the reference's solver with that preconditioner."""
        code = {"cg_ichol": 0, "bicgstab_ilut": 1, "gmres_ilut": 2}[which]
        return self._krylov(self.lib.eigref_precond_solver_f64, A, b, x0, A.rows,
                            lambda bb, x: (code, A.rows, A.nnz, A.rowptr, A.colidx, A.vals, bb, x, int(x0 is not None),
                                           tol, max_iters, uplo, ordering, droptol, fillfactor, restart))


_port = None
_ref = None


def port() -> Port:
    global _port
    if _port is None:
        _port = Port()
    return _port


def ref() -> Ref:
    global _ref
    if _ref is None:
        _ref = Ref()
    return _ref


def ref_available() -> bool:
    try:
        ref()
        return True
    except (FileNotFoundError, OSError):
        return False

"""GPU-free access to the plan that b200s_analyze_pattern builds (partition, halo lists, SpMV tiles).

Calls b200s_plan_probe, which runs the same host code (csrc/plan.cpp) without touching CUDA, so the host logic of the
multi-GPU path can be tested on a CPU-only machine (tests/test_plan.py, tests/test_dist_gloo.py)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from ._lib import Config, Stats
from .solvers import Communicator, _as_csr, _ptr


@dataclass
class PlanView:
    ghosts: np.ndarray        # sorted global column ids received from peers
    local_colidx: np.ndarray  # owned c -> c - row0 ; ghost g -> rows + g
    send_rows: np.ndarray     # local rows sent, grouped by destination rank
    send_counts: np.ndarray
    recv_counts: np.ndarray
    stats: dict


def probe(A, comm: Optional[Communicator] = None, tile_nnz: int = 0, tile_rows: int = 0) -> PlanView:
    A = _as_csr(A)
    L = _lib.lib()
    cfg = Config()
    cfg.struct_size = C.sizeof(Config)
    cfg.world = comm.world if comm else 1
    cfg.rank = comm.rank if comm else 0
    cfg.tile_nnz, cfg.tile_rows = tile_nnz, tile_rows
    if comm:
        cfg.allgather = comm.callback
    world = cfg.world
    nnz = int(A.colidx.shape[0])
    local = np.empty(nnz, np.int32)
    cap = max(1, nnz)
    ghosts = np.empty(cap, np.int64)
    send_cap = max(1, A.rows * max(1, world))
    send_rows = np.empty(send_cap, np.int32)
    sc = np.zeros(world, np.int64)
    rc = np.zeros(world, np.int64)
    st = Stats()
    st.struct_size = C.sizeof(Stats)
    rs = comm.row_starts if comm else None
    n = L.b200s_plan_probe(C.byref(cfg), A.rows, A.cols, nnz, _ptr(A.rowptr), _ptr(A.colidx), _ptr(rs), _ptr(local),
                           _ptr(ghosts), cap, _ptr(send_rows), send_cap, _ptr(sc), _ptr(rc), C.byref(st))
    if n < 0:
        msg = L.b200s_last_error(None).decode()
        if comm and comm.errors:
            msg += f" ({comm.errors[-1]!r})"
        raise _lib.B200Error(int(n), msg)
    return PlanView(ghosts[:n].copy(), local, send_rows[:int(sc.sum())].copy(), sc, rc, st.as_dict())


def canonical_csr(A, uplo: int = 3, inner_nnz=None):
    """The single-rank device matrix analyze_pattern builds from an uncompressed and / or one-triangle input:
    returns (rowptr, colidx, src) with src[k] = index into A.vals that entry k copies (GPU-free)."""
    A = _as_csr(A)
    L = _lib.lib()
    inz = None if inner_nnz is None else np.ascontiguousarray(inner_nnz, np.int32)
    nnz_in = int(A.colidx.shape[0])
    cap = 2 * nnz_in + 1
    rowptr = np.zeros(A.rows + 1, np.int32)
    colidx = np.zeros(cap, np.int32)
    src = np.zeros(cap, np.int32)
    n = L.b200s_plan_probe_csr(A.rows, nnz_in, _ptr(A.rowptr), _ptr(A.colidx), _ptr(inz), uplo, _ptr(rowptr),
                               _ptr(colidx), _ptr(src), cap)
    if n < 0:
        raise _lib.B200Error(int(n), L.b200s_last_error(None).decode())
    return rowptr, colidx[:n].copy(), src[:n].copy()


def selfadjoint_rows(A, uplo: int, comm: Optional[Communicator] = None, inner_nnz=None):
    """This rank's rows of the full self-adjoint matrix the device holds when ``A`` (this rank's row block, global
    columns) stores one triangle: returns (rowptr, global column ids, values).  With a communicator the mirror images
    stored by other ranks are fetched through its allgather -- a collective, every rank calls it (GPU-free)."""
    A = _as_csr(A)
    L = _lib.lib()
    cfg = Config()
    cfg.struct_size = C.sizeof(Config)
    cfg.world = comm.world if comm else 1
    cfg.rank = comm.rank if comm else 0
    if comm:
        cfg.allgather = comm.callback
    inz = None if inner_nnz is None else np.ascontiguousarray(inner_nnz, np.int32)
    rs = comm.row_starts if comm else None
    nnz_in = int(A.colidx.shape[0])
    vals = np.ascontiguousarray(A.vals, np.float64)
    rowptr = np.zeros(A.rows + 1, np.int32)
    args = (C.byref(cfg), A.rows, A.cols, nnz_in, _ptr(A.rowptr), _ptr(A.colidx), _ptr(inz), uplo, _ptr(rs), _ptr(vals),
            _ptr(rowptr))
    # the mirrors arriving from other ranks are not bounded by the local input: ask for the size first
    n = L.b200s_plan_probe_selfadjoint(*args, None, None, 0)
    if n < 0:
        raise _lib.B200Error(int(n), L.b200s_last_error(None).decode())
    cols = np.zeros(max(1, n), np.int64)
    out = np.zeros(max(1, n), np.float64)
    n = L.b200s_plan_probe_selfadjoint(*args, _ptr(cols), _ptr(out), int(n))
    if n < 0:
        raise _lib.B200Error(int(n), L.b200s_last_error(None).decode())
    return rowptr, cols[:n].copy(), out[:n].copy()

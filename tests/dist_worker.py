"""Worker for the multi-process tests (launched once per rank by tests/test_dist_gloo.py and tests/test_gpu_multi.py).

mode "plan" (CPU, gloo): builds the row-block plan through the GPU-free probe, performs the halo exchange the plan
prescribes with torch.distributed, multiplies the local block with numpy and checks it against the oracle's global
SpMV -- i.e. partition + ghost numbering + send lists are exactly what a correct distributed product needs.

mode "gpu" (one process per GPU): SpMV, CG and BiCGSTAB through the C ABI on the partitioned problem (a few dozen SpMV
tiles per rank, interior and boundary); every rank gathers the pieces and compares with the oracle run on the whole
problem.  CG is repeated with one stored triangle per rank (UpLo = Lower, Upper) and must give identical bits.

mode "tri" (CPU, gloo): the row-partitioned self-adjoint expansion (pattern + values of the mirrors exchanged through
the allgather callback) equals the rows of the full symmetric matrix, entry for entry.

mode "fullsize" (one process per GPU): BASELINE.json's configurations at full size -- name is cg_256, bicg_256 or
cg_512 -- row-partitioned over the visible GPUs and compared with the reference fixtures of
tests/golden/fullsize_v1.npz (tests/golden/make_fullsize.py) exactly as tests/test_gpu_fullsize.py does on one GPU.

mode "deadpeer" (one process per GPU): rank 1 never calls solve(); the others must come back with B200S_ERR_COMM
(status exchange before the launch), not hang.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make(name, big=False):
    from eigen_git_mirror_b200 import workloads as wl
    n = 48 if big else 12
    if name == "poisson3d":
        return wl.poisson3d(n), n * n
    if name == "convdiff3d":
        return wl.convdiff3d(n), n * n
    if name == "varcoef3d":
        return wl.varcoef3d(40 if big else 10), 1
    if name == "powerlaw":
        A = wl.powerlaw(60000 if big else 1500, 6, seed=5)
        # make it diagonally dominant so that CG/BiCGSTAB behave: A <- A + A^T + 40 I
        S = A.to_scipy()
        import scipy.sparse as sp
        S = (S + S.T + 40.0 * sp.identity(A.rows)).tocsr()
        S.sort_indices()
        return wl.CsrMatrix(A.rows, A.cols, S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data), 1
    raise SystemExit(name)


def block_of(A, r0, r1):
    from eigen_git_mirror_b200.workloads import CsrMatrix
    lo, hi = A.rowptr[r0], A.rowptr[r1]
    return CsrMatrix(r1 - r0, A.cols, (A.rowptr[r0:r1 + 1] - lo).astype(np.int32), A.colidx[lo:hi].copy(),
                     A.vals[lo:hi].copy(), r0)


def gather_vec(dist, local, starts, group=None):
    import torch
    world = dist.get_world_size()
    n = int(max(np.diff(starts)))
    buf = torch.zeros(n, dtype=torch.float64)
    buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    out = [torch.zeros(n, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return np.concatenate([out[q].numpy()[: int(starts[q + 1] - starts[q])] for q in range(world)])


def main():
    mode, name = sys.argv[1], sys.argv[2]
    import torch
    import torch.distributed as dist
    import eigen_git_mirror_b200 as egm
    from eigen_git_mirror_b200 import planning, workloads as wl
    from oracle import loader

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if mode == "gpu":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group(backend="gloo")
    if mode == "fullsize":
        return fullsize(name, rank, world, dist, egm, wl)
    A, align = make(name, big=(mode not in ("plan", "tri")))
    # stencils: plane-aligned equal blocks; irregular rows: blocks balanced by the bytes an iteration streams
    starts = egm.partition_rows(A.rows, world, align=align, rowptr=(A.rowptr if name == "powerlaw" else None))
    r0, r1 = int(starts[rank]), int(starts[rank + 1])
    Ab = block_of(A, r0, r1)
    comm = egm.Communicator.from_torch(starts)
    port = loader.port()
    x = wl.random_vector(A.cols, 54321)
    y_ref = port.spmv(A, x)

    if mode == "plan":
        v = planning.probe(Ab, comm)
        # 1. ghosts are exactly the external columns this block references
        ext = np.unique(Ab.colidx[(Ab.colidx < r0) | (Ab.colidx >= r1)])
        assert np.array_equal(v.ghosts, ext)
        # 2. halo exchange as prescribed: I send x[r0 + send_rows], grouped by destination, and receive my ghosts
        #    ordered by owner = ordered by global column
        send = x[r0 + v.send_rows]
        inp = [torch.from_numpy(np.ascontiguousarray(s)) for s in np.split(send, np.cumsum(v.send_counts)[:-1])]
        # gloo has no all_to_all on every build: emulate with all_gather of padded buffers
        m = int(max(1, A.rows))
        pad = torch.zeros(world, m, dtype=torch.float64)
        for q in range(world):
            pad[q, : inp[q].shape[0]] = inp[q]
        gathered = [torch.zeros(world, m, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, pad)
        ghost_vals = np.concatenate([gathered[q][rank, : int(v.recv_counts[q])].numpy() for q in range(world)])
        assert np.array_equal(ghost_vals, x[v.ghosts])
        # 3. local product on [owned | ghost] with the remapped columns equals the oracle's rows
        x_ext = np.concatenate([x[r0:r1], ghost_vals])
        Aloc = wl.CsrMatrix(Ab.rows, Ab.rows + len(v.ghosts), Ab.rowptr, v.local_colidx, Ab.vals)
        assert np.array_equal(port.spmv(Aloc, x_ext), y_ref[r0:r1])
        assert v.stats["ghosts"] == len(ext) and v.stats["halo_send"] == len(v.send_rows)
        if name == "poisson3d":  # k-slab partition: one n^2 plane per neighbour, few boundary tiles
            nb = (rank > 0) + (rank < world - 1)
            assert len(ext) == 144 * nb and v.stats["tiles_boundary"] <= 2 * nb + 1
        print(f"rank {rank}: plan ok, ghosts {len(ext)}, send {len(v.send_rows)}")
    elif mode == "tri":
        # One stored triangle, row-partitioned: every rank passes the Lower (Upper) part of ITS rows only; the plan must
        # hand each rank exactly its rows of the full self-adjoint matrix (pattern, values and summation order), with
        # the mirrors of entries stored by other ranks fetched through the allgather.
        import scipy.sparse as sp
        S = A.to_scipy()
        S = ((S + S.T) * 0.5).tocsr()        # symmetric whatever the workload was
        S.sort_indices()
        for uplo, tri in ((egm.Lower, sp.tril), (egm.Upper, sp.triu)):
            T = tri(S, format="csr")
            T.sort_indices()
            Tb = T[r0:r1]
            blk = wl.CsrMatrix(r1 - r0, A.cols, Tb.indptr.astype(np.int32), Tb.indices.astype(np.int32),
                               Tb.data.copy(), r0)
            rp, cols, vals = planning.selfadjoint_rows(blk, uplo, comm)
            want = S[r0:r1]
            assert np.array_equal(rp, want.indptr), "row pointers of the expanded block"
            assert np.array_equal(cols, want.indices), "columns (sorted input stays sorted)"
            assert np.array_equal(vals, want.data), "values, own and imported"
            # uncompressed storage with a stray entry of the other triangle (ignored, as selfadjointView does)
            if uplo == egm.Lower and blk.rows:
                pad = 2
                rp_u = (blk.rowptr + pad * np.arange(blk.rows + 1)).astype(np.int32)
                ci_u = np.zeros(int(rp_u[-1]), np.int32)
                va_u = np.full(int(rp_u[-1]), 777.0)
                inner = np.diff(blk.rowptr).astype(np.int32)
                for i in range(blk.rows):
                    n_i = inner[i]
                    ci_u[rp_u[i]:rp_u[i] + n_i] = blk.colidx[blk.rowptr[i]:blk.rowptr[i + 1]]
                    va_u[rp_u[i]:rp_u[i] + n_i] = blk.vals[blk.rowptr[i]:blk.rowptr[i + 1]]
                if r0 + 1 < A.cols:      # entry (r0, cols-1) lies in the upper triangle: must be skipped
                    ci_u[rp_u[0] + inner[0]] = A.cols - 1
                    inner[0] += 1
                unc = wl.CsrMatrix(blk.rows, A.cols, rp_u, ci_u, va_u, r0)
                rp2, cols2, vals2 = planning.selfadjoint_rows(unc, uplo, comm, inner_nnz=inner)
                assert np.array_equal(rp2, rp) and np.array_equal(cols2, cols) and np.array_equal(vals2, vals)
        print(f"rank {rank}: tri ok, block nnz {len(cols)}")
    elif mode == "deadpeer":
        s = egm.ConjugateGradient(Ab, comm=comm)
        b = np.asarray(A.to_scipy() @ wl.random_vector(A.rows, 12345))
        if rank == 1:
            # this rank "fails on the host": it reports a non-zero status instead of launching
            comm._allgather((-3).to_bytes(8, "little", signed=True))
            print(f"rank {rank}: deadpeer ok (played dead)")
        else:
            try:
                s.solve(b[r0:r1])
                raise SystemExit("solve() returned although a peer never launched")
            except egm.B200Error as e:
                assert e.status == -5 and "rank 1 failed before the launch" in str(e), str(e)
                print(f"rank {rank}: deadpeer ok ({e})")
    else:
        # ---- SpMV ----
        op = egm.SparseOperator(Ab, comm=comm)
        y = op.multiply(x[r0:r1])
        import scipy.sparse as sp
        scale = (abs(A.to_scipy()) @ np.abs(x))[r0:r1]
        assert np.all(np.abs(y - y_ref[r0:r1]) <= 1e-13 * scale), np.abs(y - y_ref[r0:r1]).max()
        # ---- solvers ----
        x_true = wl.random_vector(A.rows, 12345)
        b = np.asarray(A.to_scipy() @ x_true)
        if name != "convdiff3d":
            s = egm.ConjugateGradient(Ab, comm=comm)
            s.setTolerance(1e-10)
            xl = s.solve(b[r0:r1])
            xg = gather_vec(dist, xl, starts)
            xr, itr, errr, infor = port.cg(A, b, tol=1e-10)
            rel = np.linalg.norm(xg - xr) / np.linalg.norm(xr)
            assert s.info() == infor == 0 and abs(s.iterations() - itr) <= max(1, 0.02 * itr), (s.iterations(), itr)
            assert rel < 1e-8 and s.error() <= 1e-10, rel
            # warm start + determinism across reruns
            xl2 = s.solve(b[r0:r1])
            assert np.array_equal(xl, xl2)
            xl3 = s.solveWithGuess(b[r0:r1], xl)
            assert s.iterations() == 0
            msg = f"cg iters {s.iterations()} rel {rel:.2e}"
            # ConjugateGradient<_, Lower> / <_, Upper> row-partitioned: every rank passes one triangle of its rows; the
            # device matrix is the same full block (same summation order for sorted rows) => identical bits
            if name in ("poisson3d", "powerlaw"):
                St = A.to_scipy()
                for uplo, tri in ((egm.Lower, sp.tril), (egm.Upper, sp.triu)):
                    Tb = tri(St, format="csr")[r0:r1]
                    Tb.sort_indices()
                    blk = wl.CsrMatrix(r1 - r0, A.cols, Tb.indptr.astype(np.int32), Tb.indices.astype(np.int32),
                                       Tb.data.copy(), r0)
                    st = egm.ConjugateGradient(blk, uplo=uplo, comm=comm)
                    st.setTolerance(1e-10)
                    xt = st.solve(b[r0:r1])
                    assert st.iterations() == s.iterations() and np.array_equal(xt, xl), (uplo, st.iterations())
                    st.close()
                msg += "; Lower/Upper identical"
        else:
            msg = ""
        s2 = egm.BiCGSTAB(Ab, comm=comm)
        s2.setTolerance(1e-10)
        xl = s2.solve(b[r0:r1])
        xg = gather_vec(dist, xl, starts)
        xr, itr, errr, infor = port.bicgstab(A, b, tol=1e-10)
        rel = np.linalg.norm(xg - xr) / np.linalg.norm(xr)
        assert s2.info() == infor == 0 and abs(s2.iterations() - itr) <= max(2, 0.05 * itr), (s2.iterations(), itr)
        assert rel < 1e-8 and s2.error() <= 1e-10, rel
        xl_w = s2.solveWithGuess(b[r0:r1], xl + 1e-6)
        assert s2.info() == 0
        print(f"rank {rank}: gpu ok ({name}) {msg}; bicgstab iters {s2.iterations()} rel {rel:.2e}; "
              f"ghosts {op.stats()['ghosts']}")
        op.close(); s2.close()
    dist.barrier()
    dist.destroy_process_group()


def fullsize(name, rank, world, dist, egm, wl):
    """BASELINE configs[1], [2], [4] row-partitioned over `world` GPUs against the reference fixtures."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    from conftest import Fullsize
    from test_gpu_fullsize import KS, TOL, assert_close_to_reference, check_against_reference
    fx = Fullsize()
    n = 512 if name.endswith("512") else 256
    bicg = name.startswith("bicg")
    N = n ** 3
    starts = egm.partition_rows(N, world, align=n * n)
    r0, r1 = int(starts[rank]), int(starts[rank + 1])
    A = (wl.convdiff3d if bicg else wl.poisson3d)(n, rows=(r0, r1))
    x_true = wl.random_vector(N, 12345)
    b = np.asarray(A.to_scipy() @ x_true)
    del x_true
    comm = egm.Communicator.from_torch(starts)
    s = (egm.BiCGSTAB if bicg else egm.ConjugateGradient)(A, comm=comm)
    s.setTolerance(TOL)

    def reduce_parts(p):
        t = torch.tensor(p, dtype=torch.float64)
        dist.all_reduce(t)
        return [tuple(t.tolist())]

    tags = [f"k{k}" for k in KS] + (["full"] if fx.has(name, "full/iters") else [])
    for tag in tags:
        s.setMaxIterations(int(tag[1:]) if tag != "full" else -1)
        x = s.solve(b)
        parts = reduce_parts(check_against_reference(fx, name, tag, x, s.iterations(), s.error(), s.info(), row0=r0,
                                                     err_rtol=1e-5 if (bicg and tag == "k50") else 1e-9))
        if not (bicg and tag in ("k50",)):
            assert_close_to_reference(fx, name, tag, parts)
        if rank == 0:
            print(f"{name} x{world} {tag}: iterations {s.iterations()} error {s.error():.6e} "
                  f"(reference {fx.get(name, tag + '/iters')}, {fx.get(name, tag + '/error'):.6e})", flush=True)
    if "full" in tags:  # true residual with the distributed product
        op = egm.SparseOperator(A, comm=comm)
        r = b - op.multiply(x)
        t = torch.tensor([float(r @ r), float(b @ b)], dtype=torch.float64)
        dist.all_reduce(t)
        res = float(np.sqrt(t[0] / t[1]))
        assert res <= max(1.05 * TOL, 1.05 * fx.get(name, "full/true_residual")), res
        op.close()
    s.close()
    print(f"rank {rank}: fullsize ok ({name})")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

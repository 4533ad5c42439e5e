"""Generate tests/golden/golden_v2.npz -- additions of round 2 to the golden vectors, same method as make_golden.py
(outputs of the UNMODIFIED reference, oracle/_ref, on small seeded problems; build container only):

    python tests/golden/make_golden_v2.py

* bicgstab_restart/*: systems on which the reference's BiCGSTAB really takes the re-orthogonalisation branch
  (BiCGSTAB.h:72-81).  After the first step s is orthogonal to r0 by construction and r = s - w t, so r0.r = -w r0.t:
  an operator with t.s = s^T A s = 0 gives w = 0 and rho = 0 at the head of the second iteration.  Small integer
  systems make that exact in floating point; each is replicated 8 times on the block diagonal (n a multiple of the
  AVX-512 packet, where oracle/oracle.c is pinned bit-for-bit).  The reference keeps its restart count in a local
  variable, so the count comes from the C port -- accepted only where the port reproduces the reference's x,
  iterations() and error() BIT FOR BIT at every maxIterations = 1..8 and at convergence (the port then took the
  same branches).  One case restarts twice ("reset i only on the first restart", :80).
* cg_f32/*, bicgstab_f32/*: the float instantiations (ConjugateGradient<SparseMatrix<float>> etc.) at both ISA levels.
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("workloads", os.path.join(ROOT, "eigen-git-mirror_b200", "workloads.py"))
wl = importlib.util.module_from_spec(spec)
sys.modules["workloads"] = wl
spec.loader.exec_module(wl)
from oracle import loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# (expected restarts, precond, M, b) -- found by a brute-force search over small integer systems with the port
RESTART_SYSTEMS = [
    (1, 0, [[0, -2], [3, -3]], [0, -1]),
    (1, 0, [[-2, 0, -1], [-1, -3, -1], [2, -1, -2]], [2, -2, 0]),
    (1, 1, [[2, 0, 0], [2, -3, 2], [3, 0, -2]], [-2, 0, 0]),
    (1, 1, [[-1, -3, 2], [-1, -3, 0], [3, 0, -1]], [1, 0, -2]),
    (1, 1, [[3, -1, 3, -2], [-1, -3, 0, 1], [3, -2, 2, -3], [-1, 1, 2, 2]], [0, 0, 2, 0]),
    (1, 0, [[-1, -2, 2, -2], [0, 3, 1, 2], [-1, 2, 1, 0], [0, 0, -2, -1]], [0, 0, 0, -2]),
    (2, 1, [[-1, 1, 1], [1, 0, 0], [-3, -1, 3]], [1, 0, 0]),
]


def block_system(M, b, copies=8):
    import scipy.sparse as sp
    K = sp.kron(sp.identity(copies), sp.csr_matrix(np.asarray(M, float))).tocsr()
    K.sort_indices()
    n = K.shape[0]
    name = "restart_" + "_".join(str(int(v)) for v in np.asarray(M).ravel()) + f"_x{copies}"
    return (wl.CsrMatrix(n, n, K.indptr.astype(np.int32), K.indices.astype(np.int32), K.data.astype(np.float64), 0, name),
            np.tile(np.asarray(b, float), copies))


def main():
    refs = {"v4": loader.Ref("v4"), "v3": loader.Ref("v3")}
    port = loader.port()
    cases = {}

    def put(name, **kw):
        for k, v in kw.items():
            cases[f"{name}/{k}"] = np.asarray(v)

    def put_matrix(name, A):
        key = f"mat/{A.name}" + ("/f32" if A.vals.dtype == np.float32 else "")
        if f"{key}/rows" not in cases:
            put(key, rows=A.rows, cols=A.cols, rowptr=A.rowptr, colidx=A.colidx, vals=A.vals)
        put(name, matrix=key)

    def solver_case(name, kind, A, b, x0=None, tol=1e-10, max_iters=-1, uplo=3, precond=1, **extra):
        put_matrix(name, A)
        put(name, b=b, tol=tol, max_iters=max_iters, uplo=uplo, precond=precond, kind=kind,
            has_guess=int(x0 is not None), **extra)
        if x0 is not None:
            put(name, x0=x0)
        out = {}
        for v, R in refs.items():
            if kind == "cg":
                x, it, err, info = R.cg(A, b, x0=x0, tol=tol, max_iters=max_iters, uplo=uplo, precond=precond)
            else:
                x, it, err, info = R.bicgstab(A, b, x0=x0, tol=tol, max_iters=max_iters, precond=precond)
            put(name, **{f"x_{v}": x, f"iters_{v}": it, f"error_{v}": err, f"info_{v}": info})
            out[v] = (x, it, err, info)
        return out

    # ---- BiCGSTAB restart branch --------------------------------------------------------------------------------
    for idx, (want, pre, M, b) in enumerate(RESTART_SYSTEMS):
        A, bb = block_system(M, b)
        for k in list(range(1, 9)) + [-1]:
            tag = "full" if k < 0 else f"k{k}"
            name = f"bicgstab_restart/case{idx}/{tag}"
            got = solver_case(name, "bicgstab", A, bb, tol=1e-10, max_iters=k, precond=pre)
            xp, itp, errp, infop = port.bicgstab(A, bb, tol=1e-10, max_iters=k, precond=pre, lanes=8)
            rs = port.last_restarts
            xr, itr, errr, infor = got["v4"]
            assert np.array_equal(xp, xr) and itp == itr and errp == errr and infop == infor, \
                f"{name}: the port does not reproduce the reference bit for bit; its restart count cannot be used"
            put(name, restarts=rs)
            if k < 0:
                assert rs == want and infor == 0, (name, rs, want, infor)
                print(f"{name}: n={A.rows} precond={pre} restarts={rs} iters={itr} error={errr:.3e}")

    # ---- float solvers ---------------------------------------------------------------------------------------------
    def rnd_spd(n, density, seed):
        import scipy.sparse as sp
        rng = np.random.default_rng(seed)
        Mx = sp.random(n, n, density=density, random_state=rng, data_rvs=lambda k: rng.uniform(-1, 1, k)).tocsr()
        Mx = Mx + sp.diags(rng.uniform(0.5, 1.5, n))
        S = (Mx @ Mx.T).tocsr()
        S.sort_indices()
        return wl.CsrMatrix(n, n, S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data.astype(np.float64), 0,
                            f"random_spd_{n}_s{seed}")

    for A in (wl.poisson2d(24), wl.poisson3d(10), wl.varcoef3d(10), rnd_spd(96, 0.05, 11)):
        Af = A.astype(np.float32)
        xt = wl.random_vector(A.rows, 12345, np.float32)
        b = np.asarray(Af.to_scipy() @ xt, dtype=np.float32)
        for pre in (1, 0):
            solver_case(f"cg_f32/{A.name}/pre{pre}", "cg", Af, b, tol=1e-5, precond=pre)
        solver_case(f"cg_f32/{A.name}/default_tol", "cg", Af, b, tol=-1.0)
        for k in (0, 1, 2, 5):
            solver_case(f"cg_f32/{A.name}/traj_k{k}", "cg", Af, b, tol=-1.0, max_iters=k)
        solver_case(f"cg_f32/{A.name}/zero_rhs", "cg", Af, np.zeros(A.rows, np.float32))
        solver_case(f"cg_f32/{A.name}/guess", "cg", Af, b, x0=xt + np.float32(1e-2) * wl.random_vector(A.rows, 77, np.float32),
                    tol=1e-5)
        solver_case(f"cg_f32/{A.name}/lower", "cg", Af, b, tol=1e-5, uplo=1)
    for A in (wl.convdiff3d(10), wl.convdiff3d(8, gamma=0.9), wl.varcoef3d(8)):
        Af = A.astype(np.float32)
        xt = wl.random_vector(A.rows, 12345, np.float32)
        b = np.asarray(Af.to_scipy() @ xt, dtype=np.float32)
        for pre in (1, 0):
            solver_case(f"bicgstab_f32/{A.name}/pre{pre}", "bicgstab", Af, b, tol=1e-5, precond=pre)
        for k in (0, 1, 2, 5):
            solver_case(f"bicgstab_f32/{A.name}/traj_k{k}", "bicgstab", Af, b, tol=-1.0, max_iters=k)
        solver_case(f"bicgstab_f32/{A.name}/zero_rhs", "bicgstab", Af, np.zeros(A.rows, np.float32))
        solver_case(f"bicgstab_f32/{A.name}/guess", "bicgstab", Af, b,
                    x0=xt + np.float32(1e-2) * wl.random_vector(A.rows, 77, np.float32), tol=1e-5)

    path = os.path.join(OUT, "golden_v2.npz")
    np.savez_compressed(path, **cases)
    print(f"wrote {path}: {len(cases)} arrays, {os.path.getsize(path) / 1e6:.2f} MB; reference = "
          f"{refs['v4'].build_info}")


if __name__ == "__main__":
    main()

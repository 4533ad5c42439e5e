"""Generate tests/golden/fullsize_v2.npz: the UNMODIFIED reference (oracle/_ref) on BASELINE.json's 3D Poisson 256^3 system
with ConjugateGradient<_, Lower|Upper, IncompleteCholesky<double, Lower, NaturalOrdering>> -- the
natural ordering needs no permutation in the fixture.  Stored as in fullsize_v1.npz: iterations(), error(), info(), ||x||_2
and 4,096 sampled entries of x for the converged solve (tol 1e-10) and for the fixed-k trajectories k in {1, 5, 20}.
Build container only (~15 GB RAM, ~10 min of CPU):

    python tests/golden/make_fullsize_v2.py
"""
import importlib.util
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("workloads", os.path.join(ROOT, "eigen-git-mirror_b200", "workloads.py"))
wl = importlib.util.module_from_spec(spec)
sys.modules["workloads"] = wl
spec.loader.exec_module(wl)
from oracle import loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fullsize_v2.npz")
TOL = 1e-10
KS = (1, 5, 20)
NSAMPLE = 4096


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    R = loader.Ref("v3")  # exact square roots in m_scale: the build the product's host factorization is pinned to
    A = wl.poisson3d(n)
    x_true = wl.random_vector(A.rows, 12345)
    b = wl.rhs_from_solution(A, x_true)
    idx = np.sort(np.random.default_rng(2024).choice(A.rows, NSAMPLE, replace=False)).astype(np.int64)
    name = f"cg_ichol_{n}"
    out = {f"{name}/rows": np.asarray(A.rows), f"{name}/nnz": np.asarray(A.nnz), f"{name}/sample_idx": idx,
           f"{name}/build": np.asarray(R.build_info)}
    for tag, k in [(f"k{k}", k) for k in KS] + [("full", -1)]:
        t0 = time.time()
        x, it, err, info = R.precond_solver("cg_ichol", A, b, tol=TOL, max_iters=k, uplo=3, ordering=0)
        out[f"{name}/{tag}/iters"] = np.asarray(it)
        out[f"{name}/{tag}/error"] = np.asarray(err)
        out[f"{name}/{tag}/info"] = np.asarray(info)
        out[f"{name}/{tag}/xnorm"] = np.asarray(float(np.linalg.norm(x)))
        out[f"{name}/{tag}/x_samples"] = x[idx].copy()
        if tag == "full":
            out[f"{name}/full/err_vs_true"] = np.asarray(float(np.linalg.norm(x - x_true) / np.linalg.norm(x_true)))
            out[f"{name}/full/true_residual"] = np.asarray(float(np.linalg.norm(b - A.to_scipy() @ x) / np.linalg.norm(b)))
        print(f"{name} {tag}: iters {it} error {err:.6e} info {info} ({time.time() - t0:.1f}s)", flush=True)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} ({os.path.getsize(OUT) / 1e3:.0f} kB)")


if __name__ == "__main__":
    main()

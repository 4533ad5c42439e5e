"""Generate tests/golden/fullsize_v1.npz: the UNMODIFIED reference (oracle/_ref, Eigen compiled from /root/reference)
run on BASELINE.json's own configurations at FULL size.  Build container only (needs ~25 GB RAM and ~15 min of CPU):

    python tests/golden/make_fullsize.py [case ...]       # cases: p2d_1024 cg_256 bicg_256 cg_512 (default: all)

Per case the file stores what a test can compare without holding a second copy of the solution:
iterations(), error(), info(), ||x||_2 and 4,096 sampled entries of x (indices drawn once from PCG64 seed 2024),
for the full-convergence solve (tol 1e-10) and for the fixed-k trajectories k in {1, 2, 5, 10, 50}
(setMaxIterations(k), SURVEY.md 8c-5 / BASELINE.md section 3).  512^3 stores trajectories only: a full CPU solve
there is ~45 minutes.  Inputs are exactly the ones tests and bench.py build: workloads.poisson3d / convdiff3d /
poisson2d, x_true = workloads.random_vector(N, 12345), b = A x_true via scipy (workloads.rhs_from_solution).
Existing cases in the output file are kept when only some cases are regenerated.
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

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fullsize_v1.npz")
TOL = 1e-10
KS = (1, 2, 5, 10, 50)
NSAMPLE = 4096


def sample_indices(n):
    return np.sort(np.random.default_rng(2024).choice(n, NSAMPLE, replace=False)).astype(np.int64)


def run_case(name, A, kind, full, out):
    R = loader.ref()
    threads = R.max_threads
    x_true = wl.random_vector(A.rows, 12345)
    b = wl.rhs_from_solution(A, x_true)
    idx = sample_indices(A.rows)
    fn = R.cg if kind == "cg" else R.bicgstab
    out[f"{name}/kind"] = np.asarray(kind)
    out[f"{name}/rows"] = np.asarray(A.rows)
    out[f"{name}/nnz"] = np.asarray(A.nnz)
    out[f"{name}/sample_idx"] = idx
    out[f"{name}/bnorm"] = np.asarray(float(np.linalg.norm(b)))
    out[f"{name}/build"] = np.asarray(R.build_info)
    for k in KS:
        t0 = time.time()
        x, it, err, info = fn(A, b, tol=TOL, max_iters=k, threads=threads)
        out[f"{name}/k{k}/iters"] = np.asarray(it)
        out[f"{name}/k{k}/error"] = np.asarray(err)
        out[f"{name}/k{k}/info"] = np.asarray(info)
        out[f"{name}/k{k}/xnorm"] = np.asarray(float(np.linalg.norm(x)))
        out[f"{name}/k{k}/x_samples"] = x[idx].copy()
        print(f"{name} k={k}: iters {it} error {err:.6e} ({time.time() - t0:.1f}s)", flush=True)
    if full:
        t0 = time.time()
        x, it, err, info = fn(A, b, tol=TOL, max_iters=-1, threads=threads)
        out[f"{name}/full/iters"] = np.asarray(it)
        out[f"{name}/full/error"] = np.asarray(err)
        out[f"{name}/full/info"] = np.asarray(info)
        out[f"{name}/full/xnorm"] = np.asarray(float(np.linalg.norm(x)))
        out[f"{name}/full/x_samples"] = x[idx].copy()
        out[f"{name}/full/err_vs_true"] = np.asarray(float(np.linalg.norm(x - x_true) / np.linalg.norm(x_true)))
        out[f"{name}/full/true_residual"] = np.asarray(float(np.linalg.norm(b - A.to_scipy() @ x) / np.linalg.norm(b)))
        print(f"{name} full: iters {it} error {err:.6e} info {info} ({time.time() - t0:.1f}s, {threads} threads)", flush=True)


CASES = {
    "p2d_1024": (lambda: wl.poisson2d(1024), "cg", True),
    "cg_256": (lambda: wl.poisson3d(256), "cg", True),
    "bicg_256": (lambda: wl.convdiff3d(256), "bicgstab", True),
    "cg_512": (lambda: wl.poisson3d(512), "cg", False),
}


def main():
    want = sys.argv[1:] or list(CASES)
    out = {}
    if os.path.exists(OUT):
        with np.load(OUT) as z:
            out = {k: z[k] for k in z.keys()}
    for name in want:
        gen, kind, full = CASES[name]
        for k in [k for k in out if k.startswith(name + "/")]:
            del out[k]
        run_case(name, gen(), kind, full, out)
        np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""SpMV-only probe (config 3 of BASELINE.json and the ncu target): builds one synthetic matrix, runs the device-resident
product `reps` times through the C ABI and prints one JSON line with GB/s against the algorithmic bytes of SURVEY 8d.

    python tools/spmv_probe.py --matrix poisson3d --n 256 [--dtype f64] [--impl 1|2] [--reps 20] [--check]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make(args):
    from eigen_git_mirror_b200 import workloads as wl
    dt = np.float64 if args.dtype == "f64" else np.float32
    m = args.matrix
    if m == "poisson3d":
        return wl.poisson3d(args.n, dtype=dt)
    if m == "poisson2d":
        return wl.poisson2d(args.n, dtype=dt)
    if m == "convdiff3d":
        return wl.convdiff3d(args.n, dtype=dt)
    if m == "stencil27":
        return wl.stencil27(args.n, dtype=dt)
    if m == "banded":
        return wl.banded(args.n, args.k, dtype=dt)
    if m == "powerlaw":
        return wl.powerlaw(args.n, args.k, dtype=dt)
    raise SystemExit(f"unknown matrix {m}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--matrix", default="poisson3d")
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--k", type=float, default=16)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--impl", type=int, default=0)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--tile-nnz", type=int, default=0)
    ap.add_argument("--tile-rows", type=int, default=0)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    if args.matrix == "banded":
        args.k = int(args.k)
    import torch
    import eigen_git_mirror_b200 as egm
    A = make(args)
    op = egm.SparseOperator(A, spmv_impl=args.impl, tile_nnz=args.tile_nnz, tile_rows=args.tile_rows)
    x = np.random.default_rng(54321).uniform(-1, 1, A.cols).astype(A.vals.dtype)
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty(A.rows, dtype=xd.dtype, device="cuda")
    op.multiply_device(xd, yd, reps=3)
    best = min(op.multiply_device(xd, yd, reps=args.reps) for _ in range(3))
    st = op.stats()
    peak = 6553.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    gbs = A.spmv_bytes() / (best * 1e-3) / 1e9
    out = {"matrix": A.name, "dtype": args.dtype, "rows": A.rows, "nnz": A.nnz, "mean_nnz_row": A.nnz / max(1, A.rows),
           "impl": "staged" if st["spmv_grid"] and args.impl != 2 else "direct", "ms": best, "gbs": gbs,
           "frac_of_measured_hbm": gbs / peak, "tiles": st["tiles"], "tiles_by_lanes": st["tiles_by_lanes"],
           "tiles_stream": st["tiles_stream"], "tiles_long": st["tiles_long"], "grid": st["spmv_grid"],
           "stages": st["spmv_stages"], "smem": st["spmv_smem_bytes"]}
    if args.check:
        ref = A.to_scipy() @ x.astype(np.float64)
        y = yd.cpu().numpy().astype(np.float64)
        scale = np.abs(A.to_scipy()) @ np.abs(x.astype(np.float64)) + 1e-300
        out["max_scaled_err"] = float(np.max(np.abs(y - ref) / scale))
    print(json.dumps(out))


if __name__ == "__main__":
    main()

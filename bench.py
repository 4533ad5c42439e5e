#!/usr/bin/env python
"""bench.py -- headline benchmark of the sparse iterative-solve path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--grid 256] [--solver cg|bicgstab]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): 3D 7-point Poisson 256^3 (16.7M unknowns, 117M nnz), ConjugateGradient<double>,
Lower|Upper, Jacobi preconditioner, tol 1e-10, b = A*x_true with x_true ~ U(-1,1) -- configs[1] of BASELINE.json, the
configuration the metric is quoted on; with --gpus N the SAME problem is row-partitioned over N GPUs (strong scaling).

A "step" is one full CG solve.  `value` = CG iterations per second with b and x resident in HBM (device timing: CUDA
events on the library's stream around the graph launch, max over ranks); `e2e` = the same metric through the public
host API (pinned host buffers; H2D of b and D2H of x inside the timed region, wall clock).  `roofline` describes the
dominant kernel (the SpMV fused with p.Ap), timed alone with CUDA events on its launching stream.  `cpu_baseline` is
the unmodified reference (Eigen, oracle/_ref) on the host cores for a bounded number of iterations of the same solve.
`--impl reference` runs only that CPU arm.  Nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cg_iterations_per_sec"
UNIT = "iterations/s"
TOL = 1e-10


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def workload_name(n, solver):
    kind = "Poisson" if solver == "cg" else "convection-diffusion (gamma=0.5)"
    return (f"3D 7-point {kind} {n}^3, {'ConjugateGradient' if solver == 'cg' else 'BiCGSTAB'}<double> "
            f"+ DiagonalPreconditioner, tol {TOL:g}, b=A*x_true")


def build_block(n, solver, r0, r1):
    from eigen_git_mirror_b200 import workloads as wl
    gen = wl.poisson3d if solver == "cg" else wl.convdiff3d
    return gen(n, rows=(r0, r1))


def iteration_bytes(nnz, rows, solver):
    """Algorithmic bytes per iteration: CG 12 nnz + 4(N+1) + 13*8 N; BiCGSTAB 2(...) + 25*8 N."""
    m = 12 * nnz + 4 * (rows + 1)
    return m + 104 * rows if solver == "cg" else 2 * m + 200 * rows


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(n, solver, iters_cap, threads=None):
    """The unmodified reference on the host cores: one CG/BiCGSTAB run capped at `iters_cap` iterations."""
    from oracle import loader
    from eigen_git_mirror_b200 import workloads as wl
    R = loader.ref()
    threads = threads or R.max_threads
    A = build_block(n, solver, 0, n ** 3)
    x_true = wl.random_vector(A.rows, 12345)
    b = wl.rhs_from_solution(A, x_true)
    return R, A, b, threads


def run_reference(args):
    """--impl reference: Eigen's own CPU implementation of the path, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n, solver = args.grid, args.solver
    # bounded sample: m iterations per step so that the run ends within minutes (256^3: ~0.2 s/iteration on 8 threads)
    m = args.ref_iters or max(2, int(round(20 * (256 / n) ** 3)))
    R, A, b, threads = cpu_reference_run(n, solver, m)
    fn = R.cg if solver == "cg" else R.bicgstab
    for _ in range(args.warmup):
        fn(A, b, tol=TOL, max_iters=min(m, 2), threads=threads)
    t_total, it_total = 0.0, 0
    for _ in range(args.steps):
        _, it, _, _ = fn(A, b, tol=TOL, max_iters=m, threads=threads)
        t_total += R.last_solve_seconds
        it_total += it
    value = it_total / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, solver), "grid": n, "unknowns": A.rows, "nnz": A.nnz,
                   "sample": f"{m} iterations per step (setMaxIterations({m})), x0 = 0"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": f"{args.steps} x {m} iterations of the same solve; {R.build_info}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    import eigen_git_mirror_b200 as egm
    from eigen_git_mirror_b200 import workloads as wl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local_rank)
    if egm.device_count() < 1:
        raise SystemExit("no sm_100 device: the product has no CPU fallback")
    comm, gloo = None, None
    n, solver = args.grid, args.solver
    N = n ** 3
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
        gloo = dist.new_group(backend="gloo")
        starts = egm.partition_rows(N, world, align=n * n)  # k-slabs: one n^2 plane of halo per neighbour
        comm = egm.Communicator.from_torch(starts, group=gloo)
        r0, r1 = int(starts[rank]), int(starts[rank + 1])
    else:
        r0, r1 = 0, N

    t_setup = time.perf_counter()
    A = build_block(n, solver, r0, r1)
    x_true = wl.random_vector(N, 12345)
    b_host = np.asarray(A.to_scipy() @ x_true)           # this rank's block of b = A x_true (setup, untimed)
    nnz_global, rows_global = (7 * N - 6 * n * n), N
    Solver = egm.ConjugateGradient if solver == "cg" else egm.BiCGSTAB
    s = Solver(comm=comm, device=local_rank)
    s.compute(A)
    s.setTolerance(TOL)
    t_setup = time.perf_counter() - t_setup
    stats = s.stats()

    rows = A.rows
    b_pin = torch.empty(rows, dtype=torch.float64).pin_memory()
    b_pin.numpy()[:] = b_host
    x_pin = torch.empty(rows, dtype=torch.float64).pin_memory()
    b_dev = b_pin.cuda(non_blocking=False)
    x_dev = torch.zeros(rows, dtype=torch.float64, device="cuda")

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        import torch.distributed as dist
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm: `value` ----
    for _ in range(args.warmup):
        s.solve_device(b_dev, x_dev)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms, launches, iters_total = 0.0, 0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s.solve_device(b_dev, x_dev)
        st = s.stats()
        dev_ms += st["last_solve_ms"]
        launches += st["last_kernel_launches"]
        iters_total += s.iterations()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop()
    dev_ms = max_over_ranks(dev_ms)
    iters = s.iterations()
    err, info = s.error(), s.info()
    value = iters_total / (dev_ms * 1e-3)
    timeline = s.timeline()
    if os.environ.get("B200S_BENCH_TIMELINES"):
        print(f"[rank {rank}] timeline {json.dumps(timeline)}", file=sys.stderr)

    # ---- end-to-end arm through the host API: pinned host b -> device, solve, x -> pinned host ----
    xh = x_pin.numpy()
    bh = b_pin.numpy()
    import ctypes as C
    L = s._hd.L
    fn = L.b200s_cg_solve_f64 if solver == "cg" else L.b200s_bicgstab_solve_f64
    it_c, err_c, info_c = C.c_int64(0), C.c_double(0), C.c_int(0)

    def host_solve():
        s._hd.check(fn(s._hd.h, C.c_void_p(bh.ctypes.data), C.c_void_p(xh.ctypes.data), 0, TOL, s.maxIterations(),
                       C.byref(it_c), C.byref(err_c), C.byref(info_c)))
        return it_c.value

    for _ in range(min(args.warmup, 2)):
        host_solve()
    barrier()
    t0 = time.perf_counter()
    e2e_iters = 0
    for _ in range(args.steps):
        e2e_iters += host_solve()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = e2e_iters / e2e_s
    true_res = None

    # ---- dominant kernel alone: SpMV (+ fused dot in the solver), CUDA events on the library stream ----
    op = egm.SparseOperator(comm=comm, device=local_rank)
    op.compute(A)
    xv = torch.from_numpy(x_true[r0:r1].copy()).cuda() if world > 1 else torch.from_numpy(x_true).cuda()
    yv = torch.empty(rows, dtype=torch.float64, device="cuda")
    op.multiply_device(xv, yv, reps=5)
    barrier()
    spmv_ms = max_over_ranks(op.multiply_device(xv, yv, reps=args.spmv_reps))
    spmv_bytes = 12 * nnz_global + 4 * (rows_global + 1) + 16 * rows_global
    peak, peak_src = measured_peak_gbs()
    achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9
    it_bytes = iteration_bytes(nnz_global, rows_global, solver)
    it_gbs = it_bytes * iters_total / (dev_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"poisson3d_{n}_bytes_per_launch")
        except Exception:
            traffic = None

    # ---- parity guard: the solve converged and the true residual is below tol (checked on rank 0 at N=1) ----
    if world == 1 and not args.skip_check:
        s.solve_device(b_dev, x_dev)
        r_dev = torch.empty_like(x_dev)
        op.multiply_device(x_dev, r_dev, reps=1)
        true_res = float(torch.linalg.norm(b_dev - r_dev) / torch.linalg.norm(b_dev))

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            m = args.ref_iters or max(4, int(round(60 * (256 / n) ** 3)))
            R, Af, bf, threads = cpu_reference_run(n, solver, m)
            fnr = R.cg if solver == "cg" else R.bicgstab
            _, itc, _, _ = fnr(Af, bf, tol=TOL, max_iters=m, threads=threads)
            cpu_baseline = {"value": itc / R.last_solve_seconds, "unit": UNIT, "cores": threads, "kind": "reference",
                            "sample": f"{itc} iterations of the same solve (setMaxIterations({m})); {R.build_info}"}
        except Exception as e:  # the checker is optional at run time; say so rather than fail the bench
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {e!r}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n, solver), "grid": n, "unknowns": rows_global, "nnz": nnz_global,
                       "parallelism": f"row-block x{world}" if world > 1 else "single GPU",
                       "l2": "working set per iteration (3.2 GB at 256^3) exceeds L2; no explicit flush",
                       "iterations_per_solve": iters, "error": err, "info": info, "true_residual": true_res,
                       "loop_mode": stats["loop_mode"], "evict_first": stats["evict_first"], "spmv_grid": stats["spmv_grid"],
                       "spmv_stages": stats["spmv_stages"], "setup_s": round(t_setup, 2)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(rows * 8) * world,
                    "d2h_bytes_per_step": int(rows * 8) * world, "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(launches),
            "wall_ms_per_step": wall_ms / args.steps,
            "roofline": {"bound": "hbm", "kernel": "spmv_staged_kernel<double> (CSR SpMV, y = A p)",
                         "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                         "traffic": traffic, "peak_source": peak_src, "bytes_per_launch": spmv_bytes,
                         "ms_per_launch": spmv_ms},
            "spmv": {"gbs": achieved, "frac_of_hbm": achieved / (peak * world), "ms": spmv_ms},
            "iteration": {"bytes": it_bytes, "gbs": it_gbs, "frac_of_hbm": it_gbs / (peak * world),
                          "us_per_iteration": 1e3 * dev_ms / max(1, iters_total)},
            "timeline_rank0": timeline,
            "clocks": clocks,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line))
    s.close()
    op.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--solver", default="cg", choices=["cg", "bicgstab"])
    ap.add_argument("--spmv-reps", type=int, default=50)
    ap.add_argument("--ref-iters", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-check", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

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
the unmodified reference (Eigen, oracle/_ref) on the host cores for a bounded number of iterations of the same solve;
its x after those iterations is compared with the GPU's x at the same maxIterations (`config.parity_k_rel`).
`--impl reference` runs only that CPU arm.  Nothing here reads /root/reference at run time.

The headline line also carries `configs`: the other BASELINE.json configurations measured in the same run with the same
protocol (extra keys; the headline keys are unchanged) -- at N=1 configs[2] (BiCGSTAB 256^3), configs[0] (2D 1024^2) and
configs[3] (the SpMV sweep, see --sweep-rows) and IncompleteCholesky-preconditioned CG at 128^3; at N=8
configs[4] (512^3 CG).  --no-extras skips them.
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


def host_threads():
    """Cores this process may run on -- not OMP_NUM_THREADS, which torch.distributed.run pins to 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


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


def cpu_reference_problem(n, solver):
    """Inputs of the reference arm: the whole matrix and b = A x_true, as tests and the GPU arm build them."""
    os.environ["OMP_NUM_THREADS"] = str(host_threads())  # before libgomp loads with the checker
    from oracle import loader
    from eigen_git_mirror_b200 import workloads as wl
    R = loader.ref()
    A = build_block(n, solver, 0, n ** 3)
    x_true = wl.random_vector(A.rows, 12345)
    b = wl.rhs_from_solution(A, x_true)
    return R, A, b, host_threads()


def run_reference(args):
    """--impl reference: Eigen's own CPU implementation of the path, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n, solver = args.grid, args.solver
    # bounded sample: m iterations per step so that the run ends within minutes (256^3: ~0.2 s/iteration on 8 threads)
    m = args.ref_iters or max(2, int(round(20 * (256 / n) ** 3)))
    R, A, b, threads = cpu_reference_problem(n, solver)
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


class Env:
    """Process-group plumbing of one bench process (one per GPU)."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one process per GPU)")
        torch.cuda.set_device(self.local_rank)
        self.dist, self.gloo = None, None
        if self.world > 1:
            import torch.distributed as dist
            if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's banner off stdout: rank 0 prints ONE JSON line
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", self.local_rank))
            self.gloo = dist.new_group(backend="gloo")
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if not self.dist:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.dist:
            self.dist.all_reduce(t)
        return [float(v) for v in t.tolist()]


def make_problem(env, egm, n, solver, dims=3):
    """This rank's row block of the stencil problem, its b block and a configured solver."""
    from eigen_git_mirror_b200 import workloads as wl
    N = n ** dims
    comm = None
    if env.world > 1:
        starts = egm.partition_rows(N, env.world, align=n ** (dims - 1))  # slabs: one plane of halo per neighbour
        comm = egm.Communicator.from_torch(starts, group=env.gloo)
        r0, r1 = int(starts[env.rank]), int(starts[env.rank + 1])
    else:
        r0, r1 = 0, N
    t0 = time.perf_counter()
    A = wl.poisson2d(n, rows=(r0, r1)) if dims == 2 else build_block(n, solver, r0, r1)
    x_true = wl.random_vector(N, 12345)
    b_host = np.asarray(A.to_scipy() @ x_true)           # this rank's block of b = A x_true (setup, untimed)
    Solver = egm.ConjugateGradient if solver == "cg" else egm.BiCGSTAB
    s = Solver(comm=comm, device=env.local_rank)
    s.compute(A)
    s.setTolerance(TOL)
    nnz_global = (7 * N - 6 * n * n) if dims == 3 else (5 * N - 4 * n)
    return dict(A=A, b_host=b_host, x_true=x_true, s=s, comm=comm, r0=r0, r1=r1, N=N, nnz=nnz_global,
                setup_s=time.perf_counter() - t0)


def time_device_solves(env, s, b_dev, x_dev, steps, warmup):
    for _ in range(warmup):
        s.solve_device(b_dev, x_dev)
    env.barrier()
    dev_ms, launches, iters_total = 0.0, 0, 0
    t0 = time.perf_counter()
    for _ in range(steps):
        s.solve_device(b_dev, x_dev)
        st = s.stats()
        dev_ms += st["last_solve_ms"]
        launches += st["last_kernel_launches"]
        iters_total += s.iterations()
    env.barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    return env.max_over_ranks(dev_ms), launches, iters_total, wall_ms


def true_residual(env, op, b_dev, x_dev):
    """||b - A x|| / ||b|| with the (row-partitioned) device product and norms summed over ranks."""
    torch = env.torch
    r_dev = torch.empty_like(x_dev)
    op.multiply_device(x_dev, r_dev, reps=1)
    d = b_dev - r_dev
    num, den = env.sum_over_ranks([float(torch.dot(d, d)), float(torch.dot(b_dev, b_dev))])
    return float(np.sqrt(num / den))


def side_config(env, egm, n, solver, dims, steps, warmup, peak):
    """One of the non-headline BASELINE configurations, same protocol: device-resident solves, CUDA-event time."""
    torch = env.torch
    P = make_problem(env, egm, n, solver, dims)
    s, rows = P["s"], P["A"].rows
    b_dev = torch.from_numpy(P["b_host"]).cuda()
    x_dev = torch.zeros(rows, dtype=torch.float64, device="cuda")
    dev_ms, launches, iters_total, _ = time_device_solves(env, s, b_dev, x_dev, steps, warmup)
    op = egm.SparseOperator(comm=P["comm"], device=env.local_rank)
    op.compute(P["A"])
    res = true_residual(env, op, b_dev, x_dev)
    it_bytes = iteration_bytes(P["nnz"], P["N"], solver)
    gbs = it_bytes * iters_total / (dev_ms * 1e-3) / 1e9
    out = {"workload": (workload_name(n, solver) if dims == 3 else
                        f"2D 5-point Poisson {n}^2, ConjugateGradient<double> + DiagonalPreconditioner, tol {TOL:g}, b=A*x_true"),
           "n_gpus": env.world, "value": iters_total / (dev_ms * 1e-3), "unit": UNIT, "steps": steps,
           "iterations_per_solve": s.iterations(), "error": s.error(), "info": s.info(), "true_residual": res,
           "us_per_iteration": 1e3 * dev_ms / max(1, iters_total), "iteration_bytes": it_bytes,
           "iteration_gbs": gbs, "frac_of_hbm": gbs / (peak * env.world), "loop_mode": s.stats()["loop_mode"],
           "gpu_launches": int(launches), "setup_s": round(P["setup_s"], 2)}
    if solver == "bicgstab":
        out["restarts"] = s.stats()["last_restarts"]
    s.close()
    op.close()
    del b_dev, x_dev
    torch.cuda.empty_cache()
    return out


def multi_rhs_config(env, egm, n, cols, steps, peak):
    """CG with `cols` right-hand sides at once (the reference solves them one after the other).
    value = column-iterations per second; every column's result is bit-identical to its single-column solve."""
    from eigen_git_mirror_b200 import workloads as wl
    torch = env.torch
    A = wl.poisson3d(n)
    S = A.to_scipy()
    B = torch.from_numpy(np.stack([np.asarray(S @ wl.random_vector(A.rows, 12345 + 7 * k)) for k in range(cols)])).cuda()
    X = torch.zeros_like(B)
    s = egm.ConjugateGradient(A, device=env.local_rank)
    s.setTolerance(TOL)
    for _ in range(2):
        s.solve_device_multi(B, X, cols)
    ms, col_iters = 0.0, 0
    for _ in range(steps):
        s.solve_device_multi(B, X, cols)
        ms += s.stats()["last_solve_ms"]
        col_iters += int(np.sum(s.column_iterations))
    N, nnz = A.rows, A.nnz
    per_pass = (12 * nnz + 4 * (N + 1) + 104 * N * cols)  # one matrix stream + 13 vector passes per column
    passes = steps * int(np.max(s.column_iterations))
    gbs = per_pass * passes / (ms * 1e-3) / 1e9
    out = {"workload": f"3D 7-point Poisson {n}^3, ConjugateGradient<double> + Jacobi, {cols} right-hand sides per solve(B)",
           "value": col_iters / (ms * 1e-3), "unit": "column-iterations/s", "columns": cols,
           "batch_width": s.multi_rhs_batch(), "iterations_per_column": s.column_iterations.tolist(),
           "infos": s.column_infos.tolist(), "ms_per_solve": ms / steps, "bytes_per_batched_iteration": per_pass,
           "gbs": gbs, "frac_of_hbm": gbs / peak, "gpu_launches": int(s.stats()["last_kernel_launches"])}
    s.close()
    del B, X
    torch.cuda.empty_cache()
    return out


def precond_config(env, egm, n, steps):
    """ConjugateGradient + IncompleteCholesky on 3D Poisson n^3.  The factor is computed once on the
    host; every iteration applies it on the GPU as level-scheduled triangular solves.  Two orderings: natural (the
    reference's NaturalOrdering instantiation: ~3n narrow dependency levels, launch-latency-bound) and multi-colour
    (b200s_ordering_multicolor: red-black here, 2 wide levels per solve, more iterations).  Reported next to
    Jacobi-preconditioned CG on the same system: iterations and device time to the same tolerance."""
    from eigen_git_mirror_b200 import workloads as wl
    torch = env.torch
    A = wl.poisson3d(n)
    b = torch.from_numpy(np.asarray(A.to_scipy() @ wl.random_vector(A.rows, 12345))).cuda()
    x = torch.zeros_like(b)
    out = {"workload": f"3D 7-point Poisson {n}^3, ConjugateGradient<double> + IncompleteCholesky<double, Lower>, tol {TOL:g}"}
    for ordering in ("natural", "multicolor"):
        t0 = time.perf_counter()
        perm, colours = (None, 0) if ordering == "natural" else egm.multicolor_ordering(A)
        pre = egm.IncompleteCholesky(uplo=egm.Lower, perm=perm)
        s = egm.ConjugateGradient(A, preconditioner=pre, device=env.local_rank)
        setup_s = time.perf_counter() - t0
        s.setTolerance(TOL)
        s.solve_device(b, x)  # warm-up
        ms = 0.0
        for _ in range(steps):
            s.solve_device(b, x)
            ms += s.stats()["last_solve_ms"]
        st = s.stats()
        res = {"iterations": int(s.iterations()), "error": s.error(), "info": int(s.info()), "ms_per_solve": ms / steps,
               "gpu_launches": int(st["last_kernel_launches"]), "host_factorization_and_setup_s": round(setup_s, 2),
               "factor_nnz": int(pre.L.b200s_factors_nnz(pre.handle())), "colours": colours,
               "levels": [len(pre.stage(w).level_ptr) - 1 for w in (0, 1)]}
        s.precondition(wl.random_vector(A.rows, 777))
        s.precondition(wl.random_vector(A.rows, 777))
        res["apply_ms"] = s.stats()["last_solve_ms"]
        res["apply_launches"] = int(s.stats()["last_kernel_launches"])
        # both factors once (12 B per entry) + x gathered per entry (8 B) + the vector in and out of each of the 4 passes
        res["apply_gbs"] = ((2 * res["factor_nnz"] * 20 + 8 * A.rows * 8) / (res["apply_ms"] * 1e-3) / 1e9
                            if res["apply_ms"] > 0 else None)
        out[ordering] = res
        s.close()
    j = egm.ConjugateGradient(A, device=env.local_rank)
    j.setTolerance(TOL)
    j.solve_device(b, x)
    jms = 0.0
    for _ in range(steps):
        j.solve_device(b, x)
        jms += j.stats()["last_solve_ms"]
    out["jacobi"] = {"iterations": int(j.iterations()), "ms_per_solve": jms / steps}
    j.close()
    del b, x
    torch.cuda.empty_cache()
    return out


def spmv_sweep(env, egm, rows, peak, reps=20):
    """configs[3]: SpMV-only sweep on synthetic CSR, float and double, device-resident x / y, best of
    3 runs of `reps` back-to-back products after 5 warm-ups; GB/s = algorithmic bytes / time."""
    from eigen_git_mirror_b200 import workloads as wl
    torch = env.torch
    # `rows` for the heavy families; the light ones get 4x the rows so that a product is not launch-bound (a banded
    # k=4 product over 2^20 rows moves 134 MB: 20 us at the HBM peak, comparable to launch ramp and tail)
    fams = [("banded_k4", lambda: wl.banded(4 * rows, 4)), ("banded_k16", lambda: wl.banded(4 * rows, 16)),
            ("banded_k50", lambda: wl.banded(rows, 50)), ("banded_k100", lambda: wl.banded(rows, 100)),
            ("stencil27_192", lambda: wl.stencil27(192)),
            ("powerlaw_m8", lambda: wl.powerlaw(4 * rows, 8)), ("powerlaw_m32", lambda: wl.powerlaw(rows, 32)),
            ("powerlaw_m100", lambda: wl.powerlaw(rows, 100)), ("powerlaw_m200", lambda: wl.powerlaw(rows, 200))]
    out = {}
    for name, gen in fams:
        A = gen()
        entry = {"rows": A.rows, "nnz": A.nnz}
        for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
            Ad = A.astype(dt)
            op = egm.SparseOperator(Ad, device=env.local_rank)
            x = torch.from_numpy(wl.random_vector(A.cols, 54321, dt)).cuda()
            y = torch.empty(A.rows, dtype=x.dtype, device="cuda")
            op.multiply_device(x, y, reps=5)
            ms = min(op.multiply_device(x, y, reps=reps) for _ in range(3))
            gbs = Ad.spmv_bytes() / (ms * 1e-3) / 1e9
            st = op.stats()
            entry[tag] = {"ms": ms, "gbs": gbs, "frac_of_hbm": gbs / peak, "bytes": Ad.spmv_bytes(),
                          "tiles_by_lanes": st["tiles_by_lanes"], "tiles_stream": st["tiles_stream"],
                          "tiles_long": st["tiles_long"]}
            op.close()
            del x, y
        out[name] = entry
        del A
        torch.cuda.empty_cache()
    return out


def run_ours(args):
    import eigen_git_mirror_b200 as egm
    from eigen_git_mirror_b200 import workloads as wl

    env = Env(args)
    torch = env.torch
    world, rank = env.world, env.rank
    if egm.device_count() < 1:
        raise SystemExit("no sm_100 device: the product has no CPU fallback")
    n, solver = args.grid, args.solver
    peak, peak_src = measured_peak_gbs()

    P = make_problem(env, egm, n, solver)
    A, s, comm, r0, r1, N = P["A"], P["s"], P["comm"], P["r0"], P["r1"], P["N"]
    nnz_global, rows_global, x_true, setup_s = P["nnz"], N, P["x_true"], P["setup_s"]
    stats = s.stats()
    rows = A.rows
    b_pin = torch.empty(rows, dtype=torch.float64).pin_memory()
    b_pin.numpy()[:] = P["b_host"]
    x_pin = torch.empty(rows, dtype=torch.float64).pin_memory()
    b_dev = b_pin.cuda(non_blocking=False)
    x_dev = torch.zeros(rows, dtype=torch.float64, device="cuda")

    # ---- device-resident arm: `value` ----
    # nvidia-smi needs a few hundred ms to produce its first sample: it is started before the warm-up solves (the same
    # workload as the timed ones) so that short timed regions -- 8 GPUs: 5 x 67 ms -- are covered too
    sampler = ClockSampler(env.local_rank)
    sampler.start()
    for _ in range(args.warmup):
        s.solve_device(b_dev, x_dev)
    env.barrier()
    dev_ms, launches, iters_total, wall_ms = time_device_solves(env, s, b_dev, x_dev, args.steps, 0)
    if len(sampler.lines) < 3:
        time.sleep(0.6)
        s.solve_device(b_dev, x_dev)
    clocks = sampler.stop()
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
    env.barrier()
    t0 = time.perf_counter()
    e2e_iters = 0
    for _ in range(args.steps):
        e2e_iters += host_solve()
    env.barrier()
    e2e_s = env.max_over_ranks(time.perf_counter() - t0)
    e2e_value = e2e_iters / e2e_s

    # ---- dominant kernel alone: SpMV (+ fused dot in the solver), CUDA events on the library stream ----
    op = egm.SparseOperator(comm=comm, device=env.local_rank)
    op.compute(A)
    xv = torch.from_numpy(x_true[r0:r1].copy()).cuda()
    yv = torch.empty(rows, dtype=torch.float64, device="cuda")
    op.multiply_device(xv, yv, reps=5)
    env.barrier()
    spmv_ms = env.max_over_ranks(op.multiply_device(xv, yv, reps=args.spmv_reps))
    spmv_bytes = 12 * nnz_global + 4 * (rows_global + 1) + 16 * rows_global
    achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9
    it_bytes = iteration_bytes(nnz_global, rows_global, solver)
    it_gbs = it_bytes * iters_total / (dev_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None  # DRAM bytes per launch from the committed ncu --set full capture (not re-measured here)
    tpath = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get(f"poisson3d_{n}_bytes_per_launch")
            traffic_src = tj.get("source") if traffic is not None else None
        except Exception:
            traffic = None

    # ---- parity guards: converged, and the true residual (distributed product, norms summed over ranks) below tol ----
    true_res = None
    if not args.skip_check:
        s.solve_device(b_dev, x_dev)
        true_res = true_residual(env, op, b_dev, x_dev)

    # ---- CPU reference leg (rank 0, N=1): timed, and its x after m iterations compared with the GPU's at the same m ----
    cpu_baseline, parity_k = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            m = args.ref_iters or max(4, int(round(60 * (256 / n) ** 3)))
            R, Af, bf, threads = cpu_reference_problem(n, solver)
            fnr = R.cg if solver == "cg" else R.bicgstab
            x_cpu, itc, err_cpu, _ = fnr(Af, bf, tol=TOL, max_iters=m, threads=threads)
            cpu_baseline = {"value": itc / R.last_solve_seconds, "unit": UNIT, "cores": threads, "kind": "reference",
                            "sample": f"{itc} iterations of the same solve (setMaxIterations({m})); {R.build_info}"}
            s.setMaxIterations(m)
            s.solve_device(b_dev, x_dev)
            x_gpu = x_dev.cpu().numpy()
            parity_k = {"k": m, "rel_x": float(np.linalg.norm(x_gpu - x_cpu) / np.linalg.norm(x_cpu)),
                        "iterations": [int(s.iterations()), int(itc)], "error": [s.error(), err_cpu]}
            s.setMaxIterations(-1)
            del Af, bf, x_cpu
        except Exception as e:  # the checker is optional at run time; say so rather than fail the bench
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {e!r}"}

    s.close()
    op.close()
    del b_dev, x_dev, xv, yv, A, P
    torch.cuda.empty_cache()

    # ---- the other BASELINE configurations, same protocol (extra keys) ----
    configs = {}
    if not args.no_extras:
        try:
            if world == 1 and solver == "cg" and n == 256:
                configs["bicgstab_256"] = side_config(env, egm, 256, "bicgstab", 3, max(2, args.steps // 4), 3, peak)
                configs["poisson2d_1024"] = side_config(env, egm, 1024, "cg", 2, max(2, args.steps // 4), 3, peak)
                configs["multi_rhs_4"] = multi_rhs_config(env, egm, 256, 4, max(2, args.steps // 4), peak)
                configs["spmv_sweep"] = spmv_sweep(env, egm, args.sweep_rows, peak)
            if world == 8 and solver == "cg" and n == 256:
                configs["poisson3d_512"] = side_config(env, egm, 512, "cg", 3, 3, 3, peak)
        except Exception as e:
            configs["error"] = repr(e)
        if world == 1 and solver == "cg" and n == 256:
            try:  # kept apart: a problem here must not hide the configurations above
                configs["ichol_cg_128"] = precond_config(env, egm, 128, max(2, args.steps // 4))
            except Exception as e:
                configs["ichol_cg_128"] = {"error": repr(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n, solver), "grid": n, "unknowns": rows_global, "nnz": nnz_global,
                       "parallelism": f"row-block x{world}" if world > 1 else "single GPU",
                       "l2": "working set per iteration (3.2 GB at 256^3) exceeds L2; no explicit flush",
                       "iterations_per_solve": iters, "error": err, "info": info, "true_residual": true_res,
                       "parity_k_rel": parity_k,
                       "loop_mode": stats["loop_mode"], "evict_first": stats["evict_first"], "spmv_grid": stats["spmv_grid"],
                       "spmv_stages": stats["spmv_stages"], "setup_s": round(setup_s, 2)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(rows * 8) * world,
                    "d2h_bytes_per_step": int(rows * 8) * world, "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(launches),
            "wall_ms_per_step": wall_ms / args.steps,
            "roofline": {"bound": "hbm", "kernel": "spmv_staged_kernel<double> (CSR SpMV, y = A p)",
                         "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "bytes_per_launch": spmv_bytes,
                         "ms_per_launch": spmv_ms},
            "spmv": {"gbs": achieved, "frac_of_hbm": achieved / (peak * world), "ms": spmv_ms},
            "iteration": {"bytes": it_bytes, "gbs": it_gbs, "frac_of_hbm": it_gbs / (peak * world),
                          "us_per_iteration": 1e3 * dev_ms / max(1, iters_total)},
            "timeline_rank0": timeline,
            "clocks": clocks,
            "cpu_baseline": cpu_baseline,
            "configs": configs,
        }
        print(json.dumps(line))
    if env.dist:
        env.dist.barrier()
        env.dist.destroy_process_group()
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
    ap.add_argument("--no-extras", action="store_true", help="skip the non-headline BASELINE configurations")
    ap.add_argument("--sweep-rows", type=int, default=1 << 20,
                    help="rows of the banded / power-law matrices of the SpMV sweep (SURVEY 8d names 2^22; the default "
                         "keeps the default run within minutes -- matrix generation on the host dominates)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

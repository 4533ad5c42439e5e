#!/bin/bash
# ncu evidence for round 1 (run under gpurun, one GPU).  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
# 1. every launch of a short CG run with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 120 --csv --log-file gpurun_out/launches_cg.csv \
    python tools/cg_probe.py --n 256 --solver cg --iters 45 --loop-mode 3 > gpurun_out/launches_cg.out 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 120 --csv --log-file gpurun_out/launches_bicg.csv \
    python tools/cg_probe.py --n 256 --solver bicgstab --iters 20 --loop-mode 3 > gpurun_out/launches_bicg.out 2>&1
# 2. full captures: the SpMV (dominant) and the two CG vector kernels
ncu --set full --clock-control none --import-source on -k regex:spmv_staged -s 4 -c 2 -o gpurun_out/prof_spmv -f \
    python tools/spmv_probe.py --matrix poisson3d --n 256 --reps 3 > gpurun_out/prof_spmv.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmv_direct -s 4 -c 1 -o gpurun_out/prof_spmv_direct -f \
    python tools/spmv_probe.py --matrix poisson3d --n 256 --reps 3 --impl 2 > gpurun_out/prof_spmv_direct.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:"cg_update|cg_direction|spmv_staged" -s 9 -c 3 -o gpurun_out/prof_cg_iter -f \
    python tools/cg_probe.py --n 256 --solver cg --iters 10 --loop-mode 3 > gpurun_out/prof_cg_iter.out 2>&1
ls -la gpurun_out

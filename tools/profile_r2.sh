#!/bin/bash
# ncu evidence for round 2 (run under gpurun, ONE GPU).  Outputs land in gpurun_out/; summaries are made in the build
# container with tools/ncu_summary.py and committed under profiles/.
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# 1. launch lists (cold-cache, serialised: compare SHARES): the stream-mode CG / BiCGSTAB iteration at 256^3, and the
#    default (persistent kernel) CG solve as the bench runs it
$NCU --metrics gpu__time_duration.sum -s 12 -c 120 --csv --log-file gpurun_out/r2_launches_cg_256_stream.csv \
    python tools/cg_probe.py --n 256 --solver cg --iters 45 --loop-mode 3 > gpurun_out/r2_launches_cg.out 2>&1
$NCU --metrics gpu__time_duration.sum -s 12 -c 120 --csv --log-file gpurun_out/r2_launches_bicg_256_stream.csv \
    python tools/cg_probe.py --n 256 --solver bicgstab --iters 20 --loop-mode 3 > gpurun_out/r2_launches_bicg.out 2>&1
$NCU --metrics gpu__time_duration.sum -c 40 --csv --log-file gpurun_out/r2_launches_cg_256_persistent.csv \
    python tools/cg_probe.py --n 256 --solver cg --iters 40 --loop-mode 4 > gpurun_out/r2_launches_cg_persistent.out 2>&1
# 2. full captures
#    the FINAL vector kernels + the SpMV with its fused dot, stream mode
$NCU --set full --import-source on -k regex:"cg_update|cg_direction|spmv_staged" -s 9 -c 3 -o gpurun_out/r2_prof_cg_iter -f \
    python tools/cg_probe.py --n 256 --solver cg --iters 10 --loop-mode 3 > gpurun_out/r2_prof_cg_iter.out 2>&1
#    the persistent kernel the bench's timed region runs (12 iterations per launch to keep the 40 replays short)
$NCU --set full --import-source on -k regex:cg_persistent -s 1 -c 1 -o gpurun_out/r2_prof_cg_persistent -f \
    python tools/cg_probe.py --n 256 --solver cg --iters 12 --loop-mode 4 > gpurun_out/r2_prof_cg_persistent.out 2>&1
#    BiCGSTAB vector kernels
$NCU --set full --import-source on -k regex:"bicg_p|bicg_s|bicg_update" -s 6 -c 3 -o gpurun_out/r2_prof_bicg_iter -f \
    python tools/cg_probe.py --n 256 --solver bicgstab --iters 6 --loop-mode 3 > gpurun_out/r2_prof_bicg_iter.out 2>&1
#    one SpMV capture per sweep family (f64) + 7-pt f32
$NCU --set full --import-source on -k regex:spmv_staged -s 4 -c 1 -o gpurun_out/r2_prof_spmv_banded16 -f \
    python tools/spmv_probe.py --matrix banded --n 4194304 --k 16 --reps 3 > gpurun_out/r2_prof_spmv_banded16.out 2>&1
$NCU --set full --import-source on -k regex:spmv_staged -s 4 -c 1 -o gpurun_out/r2_prof_spmv_stencil27 -f \
    python tools/spmv_probe.py --matrix stencil27 --n 192 --reps 3 > gpurun_out/r2_prof_spmv_stencil27.out 2>&1
$NCU --set full --import-source on -k regex:spmv_staged -s 4 -c 1 -o gpurun_out/r2_prof_spmv_powerlaw32 -f \
    python tools/spmv_probe.py --matrix powerlaw --n 1048576 --k 32 --reps 3 > gpurun_out/r2_prof_spmv_powerlaw32.out 2>&1
$NCU --set full --import-source on -k regex:spmv_staged -s 4 -c 1 -o gpurun_out/r2_prof_spmv_7pt_f32 -f \
    python tools/spmv_probe.py --matrix poisson3d --n 256 --dtype f32 --reps 3 > gpurun_out/r2_prof_spmv_7pt_f32.out 2>&1
# 3. summarise on the box (the .ncu-rep files together exceed what gpurun copies back), keep the two reports worth
#    reading in the ncu UI, drop the rest
python tools/ncu_summary.py gpurun_out/r2_prof_*.ncu-rep > gpurun_out/r2_ncu_full_summary.json 2> gpurun_out/r2_ncu_summary.err
ls -la gpurun_out/*.ncu-rep
for f in gpurun_out/r2_prof_*.ncu-rep; do
  case "$f" in *cg_persistent*|*powerlaw32*) ;; *) rm -f "$f";; esac
done
du -sh gpurun_out

#!/bin/bash
# SpMV-only sweep of BASELINE.json config 3 (+ the solver grids): one JSON line per matrix, staged kernel.
P="timeout 300 python tools/spmv_probe.py --reps 20"
for n in 64 96 128 160 192 256; do $P --matrix poisson3d --n $n; done
$P --matrix poisson2d --n 1024
$P --matrix stencil27 --n 192 --check
for dt in f64 f32; do
  for k in 4 16 50 100; do $P --matrix banded --n 4194304 --k $k --dtype $dt; done
  for m in 8 32 100; do $P --matrix powerlaw --n 1048576 --k $m --dtype $dt --check; done
  $P --matrix poisson3d --n 256 --dtype $dt
done

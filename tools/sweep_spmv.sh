#!/bin/bash
# Sweep of the staged-SpMV launch geometry on one matrix (stages x CTAs/SM x tile size); one JSON line per point.
M=${1:-poisson3d}; N=${2:-256}
for cfg in "4 2 2048 256" "3 2 2048 256" "2 4 2048 256" "2 3 2048 256" "4 4 1024 128" "2 8 1024 128" "3 5 1024 128" "6 2 1024 128" "2 2 4096 512" "2 4 1536 192" "3 3 1536 192"; do
  set -- $cfg
  B200S_SPMV_STAGES=$1 B200S_SPMV_OCC=$2 B200S_SPMV_SMEM_KB=220 timeout 120 python tools/spmv_probe.py --matrix $M --n $N --impl 1 --tile-nnz $3 --tile-rows $4 --reps 20 | sed "s/^/{\"cfg\": \"S=$1 OCC=$2 NNZ=$3 ROWS=$4\"} /"
done

#!/bin/bash
# round-2 experiment: power-law SpMV against the measured gather roofline -- tile geometry / resident CTAs
for m in 8 32; do
for tn in 0 1024 512; do
  echo "== powerlaw m=$m tile_nnz=$tn"
  python tools/spmv_probe.py --matrix powerlaw --n 1048576 --k $m --reps 20 --tile-nnz $tn | cut -c1-400
  B200S_STREAM_FACTOR=0 python tools/spmv_probe.py --matrix powerlaw --n 1048576 --k $m --reps 20 --tile-nnz $tn | cut -c1-400
done; done
echo "== powerlaw m=100"; python tools/spmv_probe.py --matrix powerlaw --n 1048576 --k 100 --reps 10 | cut -c1-400
python tools/spmv_probe.py --matrix powerlaw --n 1048576 --k 100 --reps 10 --tile-nnz 1024 | cut -c1-400

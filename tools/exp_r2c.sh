#!/bin/bash
for m in 8 32 100 200; do
  echo "== powerlaw m=$m auto"; python tools/spmv_probe.py --matrix powerlaw --n 1048576 --k $m --reps 20 --check | cut -c1-420
  echo "== powerlaw m=$m no-auto-small"; B200S_AUTO_SMALL_TILES=0 python tools/spmv_probe.py --matrix powerlaw --n 1048576 --k $m --reps 20 | cut -c1-420
  echo "== powerlaw m=$m f32"; python tools/spmv_probe.py --matrix powerlaw --n 1048576 --k $m --reps 20 --dtype f32 | cut -c1-420
done
echo "== stream factor 2"; B200S_STREAM_FACTOR=2 python tools/spmv_probe.py --matrix powerlaw --n 1048576 --k 32 --reps 20 | cut -c1-420
echo "== stream factor 1"; B200S_STREAM_FACTOR=1 python tools/spmv_probe.py --matrix powerlaw --n 1048576 --k 32 --reps 20 | cut -c1-420
python -m pytest tests/test_gpu_parity.py -q -k "spmv_golden" 2>&1 | tail -2
python -m pytest tests/test_gpu_multi.py tests/test_gpu_edge.py -q 2>&1 | tail -2

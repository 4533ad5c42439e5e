#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_fullsize.py -q -x 2>&1 | tail -2
for pdl in 0 1; do
  export B200S_PDL=$pdl
  echo "== PDL=$pdl"
  python tools/cg_probe.py --n 128 --solver cg --iters 300 --loop-mode 1 | cut -c1-420
  python tools/cg_probe.py --n 256 --solver cg --iters 300 --loop-mode 1 | cut -c1-120
done

#!/bin/bash
for pdl in 0 1; do
  export B200S_PDL=$pdl
  echo "== PDL=$pdl"
  python tools/cg_probe.py --n 128 --solver cg --iters 300 --loop-mode 1 | cut -c1-120
  python tools/cg_probe.py --n 256 --solver cg --iters 300 --loop-mode 1 | cut -c1-120
  python tools/cg_probe.py --n 256 --solver bicgstab --iters 100 --loop-mode 1 | cut -c1-120
  python tools/cg_probe.py --n 1024 --matrix poisson2d --solver cg --iters 3000 --loop-mode 1 | cut -c1-120
  python tools/multi_probe.py --n 256 --cols 4 --skip-single | cut -c1-200
done
B200S_PDL=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirhs.py tests/test_gpu_edge.py -q -x 2>&1 | tail -2

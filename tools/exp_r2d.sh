#!/bin/bash
for dt in f64 f32; do
  python tools/spmv_probe.py --matrix poisson3d --n 256 --dtype $dt --reps 30 | cut -c1-260
  python tools/spmv_probe.py --matrix banded --n 4194304 --k 16 --dtype $dt --reps 30 | cut -c1-260
  python tools/spmv_probe.py --matrix stencil27 --n 192 --dtype $dt --reps 30 | cut -c1-260
  python tools/spmv_probe.py --matrix banded --n 4194304 --k 4 --dtype $dt --reps 30 | cut -c1-260
done
python tools/cg_probe.py --n 256 --solver cg --iters 300 --loop-mode 4 | cut -c1-330
python tools/cg_probe.py --n 256 --solver cg --iters 300 --loop-mode 1 | cut -c1-330
python tools/cg_probe.py --n 128 --solver cg --iters 300 --loop-mode 4 | cut -c1-330
python tools/multi_probe.py --n 256 --cols 4 --skip-single
python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirhs.py -q -x 2>&1 | tail -2

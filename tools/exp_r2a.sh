#!/bin/bash
# round-2 experiments: L2 persistence of the CG working set at the per-GPU size of 256^3 / 8, float SpMV stage count
mkdir -p gpurun_out
python - <<'PY'
import torch
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size)
from cuda import cudart
for name in ("cudaDevAttrMaxPersistingL2CacheSize", "cudaDevAttrMaxAccessPolicyWindowSize", "cudaDevAttrL2CacheSize"):
    print(name, cudart.cudaDeviceGetAttribute(getattr(cudart.cudaDeviceAttr, name), 0))
PY
for n in 128 161; do
for mode in 4 1; do
for persist in 0 1; do
  echo "== n=$n mode=$mode persist=$persist"
  B200S_L2_PERSIST=$persist python tools/cg_probe.py --n $n --solver cg --iters 300 --loop-mode $mode
done; done; done
echo "== n=128 mode=4 persist=1 setaside sweep"
for mb in 48 64 96; do B200S_L2_PERSIST=1 B200S_L2_PERSIST_MB=$mb python tools/cg_probe.py --n 128 --solver cg --iters 300 --loop-mode 4; done
echo "== n=256 auto"; python tools/cg_probe.py --n 256 --solver cg --iters 100 --loop-mode 4
for st in 2 3 4; do
  echo "== f32 stages $st"
  B200S_SPMV_STAGES_F32=$st python tools/spmv_probe.py --matrix poisson3d --n 256 --dtype f32 --reps 30
  B200S_SPMV_STAGES_F32=$st python tools/spmv_probe.py --matrix banded --n 4194304 --k 16 --dtype f32 --reps 30
  B200S_SPMV_STAGES_F32=$st python tools/spmv_probe.py --matrix stencil27 --n 192 --dtype f32 --reps 30
done

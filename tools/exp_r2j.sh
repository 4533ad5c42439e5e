#!/bin/bash
p() { python tools/cg_probe.py --n $1 --solver cg --iters 300 --loop-mode 1 $2 | sed -E "s/.*us\/iter ([0-9.]+).*spmv_pAp_us_avg': np.float64\(([0-9.]+)\).*cg_update_us_avg': np.float64\(([0-9.]+)\).*cg_direction_us_avg': np.float64\(([0-9.]+)\).*/us\/iter \1 spmv \2 update \3 direction \4/"; }
for n in 64 128 161 203 256; do
  for c in 4 6; do echo "== n=$n vec ctas/sm $c"; B200S_VEC_CTAS_PER_SM=$c p $n; done
done
for c in 4 6; do echo "== 2D 1024 vec ctas/sm $c"; B200S_VEC_CTAS_PER_SM=$c p 1024 "--matrix poisson2d"; done

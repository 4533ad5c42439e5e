#!/bin/bash
# 8 GPUs: loop driver of the row-partitioned CG after the descriptor-prefetch change (persistent kernel vs WHILE graph)
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 4 --warmup 3 --no-extras --skip-check --grid $2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('grid', d['config']['grid'], 'loop_mode', d['config']['loop_mode'], 'it/s %.0f'%d['value'], 'us/iter %.1f'%d['iteration']['us_per_iteration'], 'frac %.3f'%d['iteration']['frac_of_hbm'], 'spmv_ms %.4f'%d['spmv']['ms'], d['timeline_rank0'])"; }
for g in 256 512; do
  for m in 4 1; do echo "== grid $g multi-mode $m"; B200S_LOOP_MODE_MULTI=$m run $((29600 + g / 8 + m)) $g; done
done

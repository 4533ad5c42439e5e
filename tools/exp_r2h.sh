#!/bin/bash
# 8 GPUs: early x update behind the all-reduce (B200S_EARLY_X) on the row-partitioned 256^3 CG
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 5 --warmup 3 --no-extras --grid $2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('grid', d['config']['grid'], 'loop_mode', d['config']['loop_mode'], 'it/s %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'us/iter %.1f'%d['iteration']['us_per_iteration'], 'frac %.3f'%d['iteration']['frac_of_hbm'], 'iters', d['config']['iterations_per_solve'], 'res', d['config']['true_residual'], d['timeline_rank0'])"; }
echo "== 256 early_x=0 pdl=0"; B200S_EARLY_X=0 B200S_PDL=0 run 29701 256
echo "== 256 early_x=0"; B200S_EARLY_X=0 run 29702 256
echo "== 256 early_x=auto(1)"; run 29703 256
echo "== 256 early_x=auto unroll 8"; B200S_BODY_UNROLL=8 run 29704 256
echo "== 512 auto"; run 29705 512

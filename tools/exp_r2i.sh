#!/bin/bash
# knobs at the per-GPU size of 256^3 / 8 (= one GPU at 128^3), WHILE graph
p() { python tools/cg_probe.py --n 128 --solver cg --iters 400 --loop-mode 1 | sed -E "s/.*us\/iter ([0-9.]+).*spmv_pAp_us_avg': np.float64\(([0-9.]+)\).*cg_update_us_avg': np.float64\(([0-9.]+)\).*cg_direction_us_avg': np.float64\(([0-9.]+)\).*/us\/iter \1 spmv \2 update \3 direction \4/"; }
echo "== default"; p
echo "== default again"; p
echo "== tile_nnz 1024"; B200S_TILE_NNZ=1024 p
echo "== tile_nnz 1536"; B200S_TILE_NNZ=1536 p
echo "== tile_nnz 1024 tile_rows 128"; B200S_TILE_NNZ=1024 B200S_TILE_ROWS=128 p
echo "== stages 3 (smem 78KB)"; B200S_SPMV_STAGES=3 B200S_SPMV_SMEM_KB=80 p
echo "== vec ctas/sm 4"; B200S_VEC_CTAS_PER_SM=4 p
echo "== vec ctas/sm 8"; B200S_VEC_CTAS_PER_SM=8 p
echo "== vec ctas/sm 3"; B200S_VEC_CTAS_PER_SM=3 p
echo "== unroll 8"; B200S_BODY_UNROLL=8 p
echo "== unroll 1"; B200S_BODY_UNROLL=1 p
echo "== evict_first 0"; B200S_EVICT_FIRST=0 p
echo "== early_x 0"; B200S_EARLY_X=0 p

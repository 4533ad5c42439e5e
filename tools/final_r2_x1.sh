#!/bin/bash
# final single-GPU evidence pass of round 2: tests, bench with every BASELINE config, ncu of the final kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/r2h_pytest.log 2>&1; tail -3 gpurun_out/r2h_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -2 gpurun_out/r2h_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench_x1.json 2> gpurun_out/r2h_bench_x1.err; echo bench_rc=$?
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_ref.json 2>/dev/null; tail -c 600 gpurun_out/r2h_bench_ref.json
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -s 12 -c 120 --csv --log-file gpurun_out/r2_launches_cg_256_stream.csv python tools/cg_probe.py --n 256 --solver cg --iters 45 --loop-mode 3 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -s 12 -c 120 --csv --log-file gpurun_out/r2_launches_bicg_256_stream.csv python tools/cg_probe.py --n 256 --solver bicgstab --iters 20 --loop-mode 3 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 60 --csv --log-file gpurun_out/r2_launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"cg_update|cg_direction|spmv_staged" -s 9 -c 3 -o gpurun_out/r2_prof_cg_iter -f python tools/cg_probe.py --n 256 --solver cg --iters 10 --loop-mode 3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"bicg_p|bicg_s|bicg_update" -s 6 -c 3 -o gpurun_out/r2_prof_bicg_iter -f python tools/cg_probe.py --n 256 --solver bicgstab --iters 6 --loop-mode 3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:cg_persistent -s 1 -c 1 -o gpurun_out/r2_prof_cg_persistent -f python tools/cg_probe.py --n 256 --solver cg --iters 12 --loop-mode 4 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:spmv_staged -s 4 -c 1 -o gpurun_out/r2_prof_spmv_banded16 -f python tools/spmv_probe.py --matrix banded --n 4194304 --k 16 --reps 3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:spmv_staged -s 4 -c 1 -o gpurun_out/r2_prof_spmv_stencil27 -f python tools/spmv_probe.py --matrix stencil27 --n 192 --reps 3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:spmv_staged -s 4 -c 1 -o gpurun_out/r2_prof_spmv_powerlaw32 -f python tools/spmv_probe.py --matrix powerlaw --n 1048576 --k 32 --reps 3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:spmv_staged -s 4 -c 1 -o gpurun_out/r2_prof_spmv_7pt_f32 -f python tools/spmv_probe.py --matrix poisson3d --n 256 --dtype f32 --reps 3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"spmm_staged|cg_update_multi|cg_direction_multi" -s 6 -c 3 -o gpurun_out/r2_prof_multi4 -f python tools/multi_probe.py --n 256 --cols 4 --iters 8 --loop-mode 3 --skip-single > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2_prof_*.ncu-rep > gpurun_out/r2_ncu_full_summary.json 2> gpurun_out/r2_ncu_summary.err
rm -f gpurun_out/r2_prof_*.ncu-rep
du -sh gpurun_out

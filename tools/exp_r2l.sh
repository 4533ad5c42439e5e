#!/bin/bash
# Last short GPU call of round 2: timing probe of the preconditioned solvers, then smoke().
mkdir -p gpurun_out
timeout 26 python -u tools/precond_probe.py 128 64 > gpurun_out/r2l_precond_timing.json 2> gpurun_out/r2l_precond_timing.err
echo "probe rc=$?"; cat gpurun_out/r2l_precond_timing.json
timeout 16 python -u -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/r2l_smoke.log

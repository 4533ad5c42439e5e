#!/bin/bash
# compute-sanitizer passes -- run under gpurun, one GPU.  Logs land in gpurun_out/ and are committed under profiles/.
#   memcheck : the whole smoke() (every loop mode, including the WHILE graph)
#   synccheck, racecheck : tools/sanitize_target.py (stream launches + the persistent cooperative kernel; the tools do
#                          not support device-side cudaGraphSetConditional, which the WHILE-graph mode uses)
set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py --smoke > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/r2_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool synccheck --num-cuda-barriers 4096 --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/r2_sanitizer_synccheck.log 2>&1
echo "synccheck rc=$?" | tee -a gpurun_out/r2_sanitizer_synccheck.log
timeout 1800 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/r2_sanitizer_racecheck.log
tail -5 gpurun_out/r2_sanitizer_*.log

#!/bin/bash
# compute-sanitizer passes over a small end-to-end run (SpMV, CG, BiCGSTAB through the C ABI) -- run under gpurun.
set -x
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python __graft_entry__.py --smoke > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"
  tail -4 gpurun_out/sanitizer_$tool.log
done

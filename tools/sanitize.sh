#!/bin/bash
# compute-sanitizer passes over a small end-to-end run (SpMV, CG, BiCGSTAB through the C ABI) -- run under gpurun.
set -x
mkdir -p gpurun_out
# synccheck does not support device-side cudaGraphSetConditional (the WHILE-graph solve dies with "unspecified launch
# failure" under the tool, memcheck is clean on the same run): it is given the stream loop mode, same kernels.
for tool in memcheck synccheck; do
  if [ $tool = synccheck ]; then export B200S_LOOP_MODE=3; extra="--num-cuda-barriers 4096"; else extra=""; fi
  timeout 900 compute-sanitizer --tool $tool $extra --error-exitcode 7 python __graft_entry__.py --smoke > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"
  tail -4 gpurun_out/sanitizer_$tool.log
done

#!/bin/bash
# One short GPU call (the round's budget was nearly spent when this was written): the GPU tests of the
# incomplete-factorization preconditioners only.  Unbuffered output straight into the log so that a cut-off call still
# leaves what ran.
mkdir -p gpurun_out
PYTHONUNBUFFERED=1 timeout 112 python -u -m pytest tests/test_gpu_precond.py -m gpu -q -p no:cacheprovider -rfE \
    > gpurun_out/r2k_precond.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_precond.log
tail -40 gpurun_out/r2k_precond.log

#!/bin/bash
set -u
mkdir -p gpurun_out
t0=$SECONDS
timeout 300 python -m pytest tests/test_gpu_ebe.py tests/test_gpu_batches.py -q -x > gpurun_out/r2l_tests.log 2>&1; tail -6 gpurun_out/r2l_tests.log; echo "tests: $((SECONDS-t0)) s"
for n in 100 63 50; do
AMARU_EBE_PATCH=0 timeout 120 python profiles/ebe_quick.py $n ebe > gpurun_out/r2l_quick_mma$n.txt 2>&1; head -c 420 gpurun_out/r2l_quick_mma$n.txt; echo; echo "quick mma $n: $((SECONDS-t0)) s"
done

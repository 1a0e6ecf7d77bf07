#!/bin/bash
set -u
mkdir -p gpurun_out
t0=$SECONDS
timeout 400 python -m pytest tests/test_gpu_ebe.py tests/test_gpu_batches.py tests/test_gpu_parity.py -q -x > gpurun_out/r2p_tests.log 2>&1; tail -4 gpurun_out/r2p_tests.log; echo "tests: $((SECONDS-t0)) s"
for n in 50 63; do
AMARU_EBE_PATCH=0 timeout 120 python profiles/ebe_quick.py $n ebe > gpurun_out/r2p_quick_mma$n.txt 2>&1; head -c 200 gpurun_out/r2p_quick_mma$n.txt; echo; echo "quick mma $n: $((SECONDS-t0)) s"
done
timeout 200 python profiles/bench_configs.py 2>/dev/null | head -1 | cut -c1-400; echo "config 1: $((SECONDS-t0)) s"

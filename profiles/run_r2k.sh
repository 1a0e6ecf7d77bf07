#!/bin/bash
set -u
mkdir -p gpurun_out
t0=$SECONDS
timeout 400 python -m pytest tests/test_gpu_ebe.py tests/test_gpu_batches.py tests/test_gpu_parity.py tests/test_gpu_configs.py -q -x -k "not full_size" > gpurun_out/r2k_tests.log 2>&1; tail -6 gpurun_out/r2k_tests.log; echo "tests: $((SECONDS-t0)) s"
AMARU_EBE_PATCH=0 timeout 120 python profiles/ebe_quick.py 100 ebe > gpurun_out/r2k_quick_mma.txt 2>&1; head -c 900 gpurun_out/r2k_quick_mma.txt; echo "quick mma: $((SECONDS-t0)) s"
timeout 120 python profiles/ebe_quick.py 100 ebe > gpurun_out/r2k_quick_patch.txt 2>&1; head -c 500 gpurun_out/r2k_quick_patch.txt; echo "quick patch: $((SECONDS-t0)) s"
AMARU_EBE_PATCH=0 timeout 120 python profiles/ebe_quick.py 63 ebe > gpurun_out/r2k_quick_mma63.txt 2>&1; head -c 300 gpurun_out/r2k_quick_mma63.txt; echo "quick mma 63^3 (= 2-GPU share): $((SECONDS-t0)) s"
AMARU_EBE_PATCH_MINPATCH=0 timeout 120 python profiles/ebe_quick.py 63 ebe > gpurun_out/r2k_quick_patch63.txt 2>&1; head -c 300 gpurun_out/r2k_quick_patch63.txt; echo "quick patch 63^3: $((SECONDS-t0)) s"
AMARU_EBE_PATCH=0 timeout 120 python profiles/ebe_quick.py 50 ebe > gpurun_out/r2k_quick_mma50.txt 2>&1; head -c 300 gpurun_out/r2k_quick_mma50.txt; echo "quick mma 50^3 (= 8-GPU share): $((SECONDS-t0)) s"

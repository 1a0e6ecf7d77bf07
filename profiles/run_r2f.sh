#!/bin/bash
# Round-2 GPU pass F (one B200): patch-form operator v2 — parity tests, kernel timings, ncu, bench line
set -u
mkdir -p gpurun_out
t0=$SECONDS
timeout 400 python -m pytest tests/test_gpu_ebe.py tests/test_gpu_batches.py -q -x > gpurun_out/r2g_ebe_tests.log 2>&1; tail -4 gpurun_out/r2g_ebe_tests.log; echo "tests: $((SECONDS-t0)) s"
timeout 120 python profiles/ebe_quick.py 100 ebe > gpurun_out/r2g_quick_patch.txt 2>&1; head -c 2000 gpurun_out/r2g_quick_patch.txt; echo "quick: $((SECONDS-t0)) s"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_ebe_patch -s 12 -c 1 -f -o gpurun_out/ebe_patch_v3 \
    python profiles/ebe_quick.py 100 ebe > gpurun_out/ncu_ebe_patch_v3.log 2>&1; echo "ncu: $? $((SECONDS-t0)) s"
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; cut -c1-1200 gpurun_out/bench_r2g.json; tail -3 gpurun_out/bench_r2g.err; echo "bench: $((SECONDS-t0)) s"

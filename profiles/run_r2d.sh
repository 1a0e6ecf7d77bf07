#!/bin/bash
# Round-2 GPU pass D (one B200): patch-form operator — parity tests, kernel timings, bench line, ncu
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ebe.py tests/test_gpu_batches.py -q -x 2>&1 | tail -15 > gpurun_out/r2d_ebe_tests.log; tail -6 gpurun_out/r2d_ebe_tests.log
timeout 300 python profiles/ebe_quick.py 100 ebe > gpurun_out/r2d_quick_patch.txt 2>&1; cat gpurun_out/r2d_quick_patch.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ebe_patch -s 12 -c 1 -f -o gpurun_out/ebe_patch_v1 \
    python profiles/ebe_quick.py 100 ebe > gpurun_out/ncu_ebe_patch_v1.log 2>&1; echo "ncu: $?"
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; cut -c1-1800 gpurun_out/bench_r2d.json; tail -3 gpurun_out/bench_r2d.err

#!/bin/bash
# Round-2 GPU pass B (one B200): DMMA operator parity + timings, whole GPU suite, bench line.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ebe.py -q -x 2>&1 | tail -8 > gpurun_out/r2b_ebe_mma.log; tail -3 gpurun_out/r2b_ebe_mma.log
python profiles/ebe_quick.py 100 ebe > gpurun_out/r2b_quick_mma.txt 2>&1; cat gpurun_out/r2b_quick_mma.txt
AMARU_EBE_MMA=0 python profiles/ebe_quick.py 100 ebe > gpurun_out/r2b_quick_dfma.txt 2>&1; cat gpurun_out/r2b_quick_dfma.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ebe_mma -s 20 -c 1 -f -o gpurun_out/ebe_mma_v1 \
    python profiles/ebe_quick.py 100 ebe > gpurun_out/ncu_ebe_mma_v1.log 2>&1; echo "ncu: $?"
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2b_pytest_gpu.log; tail -6 gpurun_out/r2b_pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; cat gpurun_out/bench_r2b.json

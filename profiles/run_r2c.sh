#!/bin/bash
# Round-2 GPU pass C (one B200): whole GPU suite, kernel timings, ncu --set full of the DMMA operator, launch list of bench.py
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2c_pytest_gpu.log; tail -6 gpurun_out/r2c_pytest_gpu.log
python profiles/ebe_quick.py 100 ebe > gpurun_out/r2c_quick_mma.txt 2>&1; cat gpurun_out/r2c_quick_mma.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ebe_mma -s 20 -c 1 -f -o gpurun_out/ebe_mma_v1 \
    python profiles/ebe_quick.py 100 ebe > gpurun_out/ncu_ebe_mma_v1.log 2>&1; echo "ncu: $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2c.csv \
    python bench.py --steps 1 --warmup 3 --size 100 --no-cpu-baseline --no-e2e --cg-maxit 60 > gpurun_out/launches_r2c.log 2>&1; echo "launch list: $?"

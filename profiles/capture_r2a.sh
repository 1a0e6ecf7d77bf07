#!/bin/bash
# Round-2 evidence capture A (one B200, under gpurun):  bash profiles/capture_r2a.sh
#   1. FP64 DFMA peak microbenchmark                                   -> gpurun_out/fp64_peak_r2.json
#   2. ncu --set full of K1 (k_assemble_K, plastic trial state), K2 (k_update), the fused CG vector kernels
#      at config-3 size (1 M HEX20)                                     -> gpurun_out/{k1,k2,cgupd}_r2.ncu-rep
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_r2a.txt
profiles/fp64_peak > gpurun_out/fp64_peak_r2.json; cat gpurun_out/fp64_peak_r2.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_K -s 8 -c 1 -f \
    -o gpurun_out/k1_r2 python profiles/prof_kernels.py --size 100 --reps 1 > gpurun_out/ncu_k1.log 2>&1
echo "ncu K1: $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_update -s 0 -c 1 -f \
    -o gpurun_out/k2_r2 python profiles/prof_kernels.py --size 100 --reps 1 > gpurun_out/ncu_k2.log 2>&1
echo "ncu K2: $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cg_update|k_cg_pupdate" -s 6 -c 2 -f \
    -o gpurun_out/cgupd_r2 python profiles/prof_kernels.py --size 100 --reps 1 > gpurun_out/ncu_cgupd.log 2>&1
echo "ncu CG vector kernels: $?"
tail -8 gpurun_out/ncu_cgupd.log

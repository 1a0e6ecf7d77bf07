#!/bin/bash
# Round-1 evidence capture (one B200, run under gpurun):  bash profiles/capture_r1.sh
#   1. ncu --set full of one k_spmv_stream2 launch at the bench size  -> gpurun_out/spmv_stream2_r1.ncu-rep
#   2. ncu launch list of the bench command (first 420 launches)       -> gpurun_out/launches_r1.csv
#   3. compute-sanitizer memcheck / racecheck / synccheck on small parity tests -> gpurun_out/sanitizer_*.log
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_stream2 -s 2 -c 1 -f \
    -o gpurun_out/spmv_stream2_r1 python profiles/prof_kernels.py --size 100 --reps 2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full: $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1
echo "ncu launches: $?"
SEL='tests/test_gpu_parity.py::test_solve_matches_direct tests/test_gpu_parity.py::test_dp_apex_branch tests/test_gpu_parity.py::test_assembly_is_deterministic'
for tool in memcheck racecheck synccheck; do
    timeout 900 compute-sanitizer --tool $tool --target-processes all --print-limit 20 \
        python -m pytest $SEL -x -q -m gpu -p no:cacheprovider > gpurun_out/sanitizer_$tool.log 2>&1
    echo "sanitizer $tool: $?"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -5
done

#!/bin/bash
# Round-2 GPU pass H (one B200): whole GPU suite + BASELINE configs end to end through solve() / solve_dynamic()
set -u
mkdir -p gpurun_out
t0=$SECONDS
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2h_pytest_gpu.log; echo "suite: $((SECONDS-t0)) s"
rm -f gpurun_out/full_configs_r2.jsonl
timeout 120 python profiles/run_full_configs.py --config 2 > gpurun_out/r2h_config2.log 2>&1; head -c 1500 gpurun_out/r2h_config2.log; echo; echo "config2: $((SECONDS-t0)) s"
timeout 500 python profiles/run_full_configs.py --config 3 > gpurun_out/r2h_config3.log 2>&1; head -c 1500 gpurun_out/r2h_config3.log; echo; echo "config3: $((SECONDS-t0)) s"
timeout 500 python profiles/run_full_configs.py --config 5 > gpurun_out/r2h_config5.log 2>&1; head -c 1500 gpurun_out/r2h_config5.log; echo; echo "config5: $((SECONDS-t0)) s"

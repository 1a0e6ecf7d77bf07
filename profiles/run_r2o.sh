#!/bin/bash
# Round-2 closing pass (one B200): whole GPU suite, every BASELINE config (one Newton iteration each), ncu launch list of the
# bench command, bench line
set -u
mkdir -p gpurun_out
t0=$SECONDS
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2o_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2o_pytest_gpu.log; echo "suite: $((SECONDS-t0)) s"
timeout 400 python profiles/bench_configs.py > gpurun_out/bench_configs_r2.jsonl 2> gpurun_out/bench_configs_r2.err; cut -c1-420 gpurun_out/bench_configs_r2.jsonl; echo "configs: $((SECONDS-t0)) s"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_r2.log 2>&1; echo "launch list: $? $((SECONDS-t0)) s"
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; cut -c1-600 gpurun_out/bench_r2_final.json; echo "bench: $((SECONDS-t0)) s"

#!/bin/bash
# 2-GPU rehearsal: group tests, config 4 through solve(ngpus=2) at reduced scale, bench N=2
set -u
export AMARU_P2P_TIMEOUT_MS=5000
mkdir -p gpurun_out
t0=$SECONDS
timeout 300 python -m pytest tests/test_gpu_group.py -q -x > gpurun_out/r2m_group_tests.log 2>&1; tail -4 gpurun_out/r2m_group_tests.log; echo "group tests: $((SECONDS-t0)) s"
timeout 200 python profiles/run_full_configs.py --config 4 --ngpus 2 --scale 4 --out gpurun_out/config4_rehearsal.jsonl 2>&1 | head -3 | cut -c1-900; echo "config4 rehearsal: $((SECONDS-t0)) s"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29720 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; cut -c1-400 gpurun_out/bench_r2_n2.json; echo "bench: $((SECONDS-t0)) s"

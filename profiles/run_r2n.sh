#!/bin/bash
# 8-GPU pass (gpurun --gpus 8): bench line of the 1 M-element config at N = 8 (fused loop), BASELINE configs[3] (TET10 5 M
# Drucker-Prager) through solve(ana, ngpus=8) on ONE handle, partitioned parity at 8 ranks
set -u
export AMARU_P2P_TIMEOUT_MS=5000
mkdir -p gpurun_out
t0=$SECONDS
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29730 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_r2_n8.json 2> gpurun_out/bench_r2_n8.err; cut -c1-1500 gpurun_out/bench_r2_n8.json; tail -2 gpurun_out/bench_r2_n8.err; echo "bench N=8: $((SECONDS-t0)) s"
timeout 420 python profiles/run_full_configs.py --config 4 --ngpus 8 --out gpurun_out/full_config4_n8_r2.jsonl > gpurun_out/r2n_config4.log 2>&1; head -c 1500 gpurun_out/r2n_config4.log; echo; echo "config 4 on 8 GPUs: $((SECONDS-t0)) s"
out=gpurun_out/mgpu_check_n8_r2.txt; : > $out
port=29740
for mode in "AMARU_P2P=0" "AMARU_P2P=1 AMARU_P2P_FUSED=1"; do
  port=$((port+1))
  echo "== HEX20 10 | $mode" >> $out
  env $mode timeout 120 $TR --master-port $port tests/mgpu_check.py HEX20 10 2>&1 | grep "operator kernel\|ranks,\|Error\|error\|Traceback" >> $out
  echo "rc=$?" >> $out
done
cat $out | cut -c1-220; echo "mgpu_check: $((SECONDS-t0)) s"

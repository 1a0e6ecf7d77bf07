#!/bin/bash
# Round-2 multi-GPU pass (run with gpurun --gpus N): partitioned parity (rank-per-GPU: NCCL, peer-memory kernels, fused loop),
# group-handle tests, bench lines.   usage: bash profiles/run_r2j.sh N
set -u
N=${1:-2}
export AMARU_P2P_TIMEOUT_MS=5000
mkdir -p gpurun_out
t0=$SECONDS
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
out=gpurun_out/mgpu_check_n${N}_r2.txt; : > $out
port=29610
for cfg in "HEX20 8" "TET10 6" "HEX8 9"; do
  for mode in "AMARU_P2P=0" "AMARU_P2P=1 AMARU_P2P_FUSED=0" "AMARU_P2P=1 AMARU_P2P_FUSED=1 AMARU_EBE_PATCH_MINFILL=0 AMARU_EBE_PATCH_MINPATCH=0" "AMARU_P2P=1 AMARU_P2P_FUSED=1"; do
    port=$((port+1))
    echo "== $cfg | $mode" >> $out
    env $mode timeout 150 $TR --master-port $port tests/mgpu_check.py $cfg 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | grep "operator kernel\|ranks,\|Error\|error\|Traceback" >> $out
    echo "rc=$?" >> $out
  done
done
cat $out | cut -c1-220; echo "mgpu_check: $((SECONDS-t0)) s"
if [ "$N" = "2" ]; then
  timeout 500 python -m pytest tests/test_gpu_group.py tests/test_gpu_multi.py -q -x > gpurun_out/r2j_group_tests_n${N}.log 2>&1; tail -5 gpurun_out/r2j_group_tests_n${N}.log; echo "group tests: $((SECONDS-t0)) s"
fi
timeout 300 $TR --master-port 29710 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_r2_n${N}_fused.json 2> gpurun_out/bench_r2_n${N}_fused.err; cut -c1-900 gpurun_out/bench_r2_n${N}_fused.json; tail -3 gpurun_out/bench_r2_n${N}_fused.err; echo "bench fused: $((SECONDS-t0)) s"

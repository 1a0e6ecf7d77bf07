#!/bin/bash
# sweep of the streamed-SpMV launch parameters (warps/CTA, pipeline stages, tile size); prints ms per SpMV+dot launch
SIZE=${1:-100}
for cfg in "8 8 256" "8 6 320" "4 8 256" "8 4 448" "8 12 160"; do
  set -- $cfg
  echo -n "warps=$1 stages=$2 tile=$3 : "
  AMARU_SPMV_WARPS=$1 AMARU_SPMV_STAGES=$2 AMARU_SPMV_TILE=$3 timeout 300 python profiles/prof_kernels.py --size $SIZE 2>&1 | grep "spmv" || echo failed
done
echo -n "simple per-lane kernel : "; AMARU_SPMV_SIMPLE=1 timeout 300 python profiles/prof_kernels.py --size $SIZE 2>&1 | grep spmv

#!/bin/bash
# sweep of the symmetric-storage SpMV (k_spmv_sym) against the full-storage kernel; prints ms per SpMV+dot launch
SIZE=${1:-100}
run() { echo -n "$* : "; env "$@" timeout 300 python profiles/prof_kernels.py --size $SIZE 2>&1 | grep -E "spmv\+dot|Error|error" | head -2 | tr '\n' ' '; echo; }
run AMARU_SPMV_SYM=0
run AMARU_SPMV_SYM=1
run AMARU_SPMV_SYM=1 AMARU_SPMV_TILE=174
run AMARU_SPMV_SYM=1 AMARU_SPMV_TILE=290
run AMARU_SPMV_SYM=1 AMARU_SPMV_TILE=174 AMARU_SPMV_STAGES=3
run AMARU_SPMV_SYM=1 AMARU_SPMV_TILE=116 AMARU_SPMV_STAGES=3
run AMARU_SPMV_SYM=1 AMARU_SPMV_XD=1

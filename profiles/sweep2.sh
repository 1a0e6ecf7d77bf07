#!/bin/bash
SIZE=${1:-100}
run() { echo -n "$* : "; env "$@" timeout 300 python profiles/prof_kernels.py --size $SIZE 2>&1 | grep "spmv" || echo failed; }
run AMARU_SPMV_SLEEP=100
run AMARU_SPMV_SLEEP=20
run AMARU_SPMV_SLEEP=300
run AMARU_SPMV_SLEEP=800

#!/bin/bash
# compute-sanitizer over the round-2 kernels (one B200): memcheck on the operator / plane-stress / axisymmetric / multi-batch
# tests, racecheck + synccheck on the patch form and the persistent colour form of the matrix-free operator
set -u
mkdir -p gpurun_out/sanitizer_r2
t0=$SECONDS
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ebe.py tests/test_gpu_parity.py -q -x -k "many_patches or mass_term or axisym or plane_stress or operator_independent" > gpurun_out/sanitizer_r2/memcheck.txt 2>&1; echo "memcheck rc=$? $((SECONDS-t0)) s"; tail -4 gpurun_out/sanitizer_r2/memcheck.txt
AMARU_EBE_PATCH_MINFILL=0 AMARU_EBE_PATCH_MINPATCH=0 timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/dbg_patch.py HEX20 6 > gpurun_out/sanitizer_r2/racecheck_patch.txt 2>&1; echo "racecheck patch rc=$? $((SECONDS-t0)) s"; tail -4 gpurun_out/sanitizer_r2/racecheck_patch.txt
AMARU_EBE_PATCH=0 timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/dbg_patch.py HEX20 6 > gpurun_out/sanitizer_r2/racecheck_colour.txt 2>&1; echo "racecheck colour rc=$? $((SECONDS-t0)) s"; tail -4 gpurun_out/sanitizer_r2/racecheck_colour.txt
AMARU_EBE_PATCH_MINFILL=0 AMARU_EBE_PATCH_MINPATCH=0 timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python profiles/dbg_patch.py QUAD8 12 > gpurun_out/sanitizer_r2/synccheck_patch.txt 2>&1; echo "synccheck rc=$? $((SECONDS-t0)) s"; tail -3 gpurun_out/sanitizer_r2/synccheck_patch.txt

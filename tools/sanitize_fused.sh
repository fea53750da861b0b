#!/bin/bash
# Run on the GPU box (under gpurun): compute-sanitizer over the fused
# small-problem path (memcheck: global / local accesses of k_solve_small and
# k_finish_trial; racecheck: the shared-memory Cholesky of the shared block).
set -u
OUT=gpurun_out
SEL='fused and (indexing_all or all_fixed_but_tau or C3_100 or C3_1500 or C2_small or C5_scaled)'
for tool in memcheck racecheck; do
    compute-sanitizer --tool $tool --error-exitcode 9 \
        python -m pytest tests/test_gpu_parity.py -q -x -k "$SEL" > $OUT/sanitizer_${tool}_fused.log 2>&1
    echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $OUT/sanitizer_${tool}_fused.log | tail -n 3
done

#!/bin/bash
# Round 2: compute-sanitizer over the kernels that are new or changed in this round -- warp-specialised,
# software-pipelined and register-accumulating Jacobian sweeps, staged chi2 sweep, shared-memory-resident
# k_solve_small, device-driven loop, block-arrow inverse, pipelined upload.
set -u
OUT=gpurun_out
mkdir -p $OUT
SEL='(fused and (indexing_all or all_fixed_but_tau or C3_100 or C3_small or C2_small or C5_small)) or block_arrow_inverse or many_tiles'
for tool in memcheck racecheck; do
    timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 \
        python -m pytest tests/test_gpu_parity.py -q -x -k "$SEL" > $OUT/r02_sanitizer_${tool}.log 2>&1
    echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $OUT/r02_sanitizer_${tool}.log | tail -n 3
done
# the staged chi2 sweep and the pipelined upload need long data sets: a scaled config-2 fit through bench.py
for tool in memcheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 \
        python bench.py --workload C2 --sets 8 --points 400000 --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-workloads \
        > $OUT/r02_sanitizer_${tool}_c2_long.log 2>&1
    echo "$tool (C2 long, e2e incl.) exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/r02_sanitizer_${tool}_c2_long.log | tail -n 2
done

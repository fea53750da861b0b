#!/bin/bash
# Round 2, GPU pass x: config-5 captures of the final build (cost table)
set -u
OUT=gpurun_out
TAG=r02z
cap() { # tag workload kernel-regex extra-args
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 \
        -o $OUT/${TAG}_prof_$1 -f python bench.py --workload $2 --steps 2 --warmup 1 --repeats 1 \
        --no-cpu-baseline --no-e2e --no-workloads $4 > $OUT/${TAG}_prof_$1.log 2>&1
    tail -1 $OUT/${TAG}_prof_$1.log
}
cap C5_k_jacobian C5 k_jacobian "--sets 25000"
cap C5_k_chi2 C5 k_chi2 "--sets 25000"

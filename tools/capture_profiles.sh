#!/bin/bash
# Run on the GPU box (under gpurun): the bench line and the ncu captures that
# tools/summarize_profiles.py turns into profiles/<tag>_*.
#   tools/capture_profiles.sh r01
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-e2e"
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_$TAG.csv $BENCH > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_jacobian -s 1 -c 1 -f \
    -o $OUT/prof_${TAG}_k_jacobian $BENCH > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_chi2 -s 2 -c 1 -f \
    -o $OUT/prof_${TAG}_k_chi2 $BENCH > /dev/null 2>&1
cat $OUT/bench_$TAG.json

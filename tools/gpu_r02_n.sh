#!/bin/bash
# Round 2, GPU pass n: wall-clock stamps of the stages of k_solve_small (developer variant builds)
set -u
OUT=gpurun_out
mkdir -p $OUT
for v in "$@"; do
GADFIT_B200_LIB=$PWD/gadfit_b200/_variants/libgadfit_b200_$v.so timeout 600 python bench.py --workload C3 --steps 6 --warmup 2 --repeats 1 --no-cpu-baseline --no-e2e --no-workloads \
      2> $OUT/r02n_solve_small_stages_$v.log > /dev/null
echo "== $v"; grep "solve_small stages" $OUT/r02n_solve_small_stages_$v.log | tail -4
done

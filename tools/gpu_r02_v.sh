#!/bin/bash
# Round 2, GPU pass v: the GPU suite, smoke(), the bench lines and the config-4 captures of the build with the
# config-4 occupancy hints
set -u
OUT=gpurun_out
TAG=r02z
mkdir -p $OUT
GADFIT_PARITY_LOG=$OUT/${TAG}_parity_log.json timeout 1500 python -m pytest tests -x -q -m gpu \
    > $OUT/${TAG}_pytest_gpu.log 2>&1
tail -n 6 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > $OUT/${TAG}_bench_C2.json 2> $OUT/${TAG}_bench_C2.err
cut -c1-300 $OUT/${TAG}_bench_C2.json
for w in C3 C4 C5; do
  timeout 900 python bench.py --workload $w --no-cpu-baseline > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err
  cut -c1-250 $OUT/${TAG}_bench_$w.json
done
cap() { # tag workload kernel-regex extra-args
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 \
        -o $OUT/${TAG}_prof_$1 -f python bench.py --workload $2 --steps 2 --warmup 1 --repeats 1 \
        --no-cpu-baseline --no-e2e --no-workloads $4 > $OUT/${TAG}_prof_$1.log 2>&1
    tail -1 $OUT/${TAG}_prof_$1.log
}
cap C4_k_jacobian C4 k_jacobian "--sets 64"
cap C4_k_chi2 C4 k_chi2 "--sets 64"
cap C4_k_omega C4 k_omega "--sets 64"

#!/bin/bash
# Round 2, final GPU pass: full GPU suite, the default bench line, the reference arm, the ncu launch
# lists and one `ncu --set full` capture per (config, sweep kernel) for profiles/ and the cost table
set -u
OUT=gpurun_out
TAG=${1:-r02z}
PART=${2:-a}   # a: suite + bench lines + launch lists + config 2/3 captures; b: config 4/5 captures (64 MiB per call)
mkdir -p $OUT
if [ "$PART" = "a" ]; then
GADFIT_PARITY_LOG=$OUT/${TAG}_parity_log.json timeout 1500 python -m pytest tests -m gpu -q \
    > $OUT/${TAG}_pytest_gpu.log 2>&1
tail -n 8 $OUT/${TAG}_pytest_gpu.log
timeout 900 python bench.py > $OUT/${TAG}_bench_C2.json 2> $OUT/${TAG}_bench_C2.err
cut -c1-600 $OUT/${TAG}_bench_C2.json
tail -3 $OUT/${TAG}_bench_C2.err
for w in C3 C4 C5; do
  timeout 900 python bench.py --workload $w --no-cpu-baseline > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err
  cut -c1-300 $OUT/${TAG}_bench_$w.json
done
timeout 1200 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
cut -c1-600 $OUT/${TAG}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches_C2.csv python bench.py --workload C2 --steps 3 --warmup 1 \
    --repeats 1 --no-cpu-baseline --no-e2e --no-workloads > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches_C3.csv python bench.py --workload C3 --steps 10 --warmup 1 \
    --repeats 1 --no-cpu-baseline --no-e2e --no-workloads > /dev/null 2>&1
fi
cap() { # tag workload kernel-regex extra-args
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 \
        -o $OUT/${TAG}_prof_$1 -f python bench.py --workload $2 --steps 2 --warmup 1 --repeats 1 \
        --no-cpu-baseline --no-e2e --no-workloads $4 > $OUT/${TAG}_prof_$1.log 2>&1
    tail -1 $OUT/${TAG}_prof_$1.log
}
if [ "$PART" = "a" ]; then
cap C2_k_jacobian C2 k_jacobian ""
cap C2_k_chi2 C2 k_chi2 ""
cap C3_k_jacobian C3 k_jacobian ""
cap C3_k_chi2 C3 k_chi2 ""
cap C3_k_solve_small C3 k_solve_small ""
else
cap C4_k_jacobian C4 k_jacobian "--sets 64"
cap C4_k_chi2 C4 k_chi2 "--sets 64"
cap C4_k_omega C4 k_omega "--sets 64"
cap C5_k_jacobian C5 k_jacobian "--sets 25000"
cap C5_k_chi2 C5 k_chi2 "--sets 25000"
fi
du -sh $OUT

#!/bin/bash
# Round 2, first GPU pass (under gpurun): the GPU test-suite, A/B of the two
# Jacobian sweeps on config 2, quick lines of the other configs, launch lists
# and `ncu --set full` captures of the sweeps of configs 2, 4 and 5.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/r02a_smi.txt 2>&1
GADFIT_PARITY_LOG=$OUT/r02a_parity_log.json timeout 1500 python -m pytest tests -m gpu -q -x \
    > $OUT/r02a_pytest_gpu.log 2>&1
tail -n 25 $OUT/r02a_pytest_gpu.log
show() {
    python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print(sys.argv[1], "unreadable", e); sys.exit(0)
print(sys.argv[1], "ms/step %.4f" % d["ms_per_step"], "launches", d["gpu_launches"],
      {k: round(v * 1e6 / max(d["phase_calls"][k], 1), 1) for k, v in d["phase_seconds"].items()},
      "K1 us %.1f frac %.3f" % (d["roofline"]["seconds_per_launch"] * 1e6, d["roofline"]["frac"] or 0),
      "e2e %.3g" % (d["e2e"]["value"] if d.get("e2e") else 0))
PY
}
timeout 600 python bench.py --workload C2 --steps 10 --warmup 3 > $OUT/r02a_bench_C2.json 2> $OUT/r02a_bench_C2.err
show $OUT/r02a_bench_C2.json
GADFIT_B200_JAC=legacy timeout 600 python bench.py --workload C2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
    > $OUT/r02a_bench_C2_legacy.json 2> $OUT/r02a_bench_C2_legacy.err
show $OUT/r02a_bench_C2_legacy.json
for w in C3 C4 C5; do
    timeout 600 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-e2e \
        > $OUT/r02a_bench_$w.json 2> $OUT/r02a_bench_$w.err
    show $OUT/r02a_bench_$w.json
done
# launch lists (cold-cache, serialised: shares only)
for w in C2 C4 C5; do
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
        --log-file $OUT/r02a_launches_$w.csv python bench.py --workload $w --steps 3 --warmup 1 \
        --repeats 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
done
# full captures of the sweeps: C2 full size; C4 / C5 at a quarter of the data sets
cap() { # tag workload kernel-regex extra-args
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 \
        -o $OUT/r02a_prof_$1 -f python bench.py --workload $2 --steps 2 --warmup 1 --repeats 1 \
        --no-cpu-baseline --no-e2e $4 > $OUT/r02a_prof_$1.log 2>&1
}
cap C2_k_jacobian C2 k_jacobian ""
cap C2_k_chi2 C2 k_chi2 ""
cap C4_k_jacobian C4 k_jacobian "--sets 64"
cap C4_k_chi2 C4 k_chi2 "--sets 64"
cap C4_k_omega C4 k_omega "--sets 64"
cap C5_k_jacobian C5 k_jacobian "--sets 25000"
cap C5_k_chi2 C5 k_chi2 "--sets 25000"
ls -la $OUT | grep r02a

#!/bin/bash
# Round 2, second GPU pass: tests + config 2/3 benches + captures of config 2
set -u
OUT=gpurun_out
mkdir -p $OUT
GADFIT_PARITY_LOG=$OUT/r02b_parity_log.json timeout 1500 python -m pytest tests -m gpu -q -x \
    > $OUT/r02b_pytest_gpu.log 2>&1
tail -n 30 $OUT/r02b_pytest_gpu.log
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
for k, v in (d.get("workloads") or {}).items():
    print("   ", k, "ms/step %.4f" % v["ms_per_step"], {a: round(b * 1e6, 1) for a, b in v["seconds_per_call"].items()},
          "frac %.3f (v1 %.3f)" % (v["roofline"]["frac"], v["roofline"]["frac_v1"]))
PY
}
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/r02b_bench_C2.json 2> $OUT/r02b_bench_C2.err
show $OUT/r02b_bench_C2.json
tail -3 $OUT/r02b_bench_C2.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/r02b_bench_reference.json 2> $OUT/r02b_bench_reference.err
cat $OUT/r02b_bench_reference.json | cut -c1-600
timeout 300 python tools/e2e_breakdown.py > $OUT/r02b_e2e_breakdown.log 2>&1
cat $OUT/r02b_e2e_breakdown.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file $OUT/r02b_launches_C2.csv python bench.py --workload C2 --steps 3 --warmup 1 \
    --repeats 1 --no-cpu-baseline --no-e2e --no-workloads > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file $OUT/r02b_launches_C3.csv python bench.py --workload C3 --steps 3 --warmup 1 \
    --repeats 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
cap() { # tag workload kernel-regex extra-args
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 \
        -o $OUT/r02b_prof_$1 -f python bench.py --workload $2 --steps 2 --warmup 1 --repeats 1 \
        --no-cpu-baseline --no-e2e --no-workloads $4 > $OUT/r02b_prof_$1.log 2>&1
}
cap C2_k_jacobian C2 k_jacobian ""
cap C2_k_chi2 C2 k_chi2 ""
cap C3_k_jacobian C3 k_jacobian ""
cap C3_k_chi2 C3 k_chi2 ""
cap C3_k_solve_small C3 k_solve_small ""
ls -la $OUT | grep r02b

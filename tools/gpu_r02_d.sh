#!/bin/bash
# Round 2, fourth GPU pass: full GPU suite, default bench line, device-driven loop A/B on config 3
set -u
OUT=gpurun_out
mkdir -p $OUT
GADFIT_PARITY_LOG=$OUT/r02d_parity_log.json timeout 1500 python -m pytest tests -m gpu -q -x \
    > $OUT/r02d_pytest_gpu.log 2>&1
tail -n 30 $OUT/r02d_pytest_gpu.log
show() {
    python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print(sys.argv[1], "unreadable", e); sys.exit(0)
print(sys.argv[1], "ms/step %.4f" % d["ms_per_step"], "launches", d["gpu_launches"],
      {k: round(v * 1e6 / max(d["phase_calls"][k], 1), 1) for k, v in d["phase_seconds"].items()}, d["phase_calls"],
      "K1 us %.1f frac %.3f" % (d["roofline"]["seconds_per_launch"] * 1e6, d["roofline"]["frac"] or 0),
      "e2e %.3g" % (d["e2e"]["value"] if d.get("e2e") else 0))
for k, v in (d.get("workloads") or {}).items():
    print("   ", k, "ms/step %.4f" % v["ms_per_step"], {a: round(b * 1e6, 1) for a, b in v["seconds_per_call"].items()},
          v["phase_calls"], "launches", v["gpu_launches"], "frac %.3f (v1 %.3f)" % (v["roofline"]["frac"], v["roofline"]["frac_v1"]))
PY
}
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/r02d_bench_C2.json 2> $OUT/r02d_bench_C2.err
show $OUT/r02d_bench_C2.json
tail -3 $OUT/r02d_bench_C2.err
for dl in 1 0; do
    GADFIT_B200_DEVICE_LOOP=$dl timeout 600 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
        > $OUT/r02d_bench_C3_dl$dl.json 2> $OUT/r02d_bench_C3_dl$dl.err
    show $OUT/r02d_bench_C3_dl$dl.json
    tail -3 $OUT/r02d_bench_C3_dl$dl.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file $OUT/r02d_launches_C3.csv python bench.py --workload C3 --steps 10 --warmup 1 \
    --repeats 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file $OUT/r02d_launches_C2.csv python bench.py --workload C2 --steps 3 --warmup 1 \
    --repeats 1 --no-cpu-baseline --no-e2e --no-workloads > /dev/null 2>&1
cap() { # tag workload kernel-regex extra-args
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 \
        -o $OUT/r02d_prof_$1 -f python bench.py --workload $2 --steps 2 --warmup 1 --repeats 1 \
        --no-cpu-baseline --no-e2e --no-workloads $4 > $OUT/r02d_prof_$1.log 2>&1
}
cap C4_k_jacobian C4 k_jacobian "--sets 64"
cap C4_k_chi2 C4 k_chi2 "--sets 64"
cap C4_k_omega C4 k_omega "--sets 64"
cap C5_k_jacobian C5 k_jacobian "--sets 25000"
cap C5_k_chi2 C5 k_chi2 "--sets 25000"
du -sh $OUT
ls -la $OUT | grep r02d

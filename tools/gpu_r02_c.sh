#!/bin/bash
# Round 2, third GPU pass: software-pipelined vs warp-specialised Jacobian sweep
set -u
OUT=gpurun_out
mkdir -p $OUT
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
for j in sp ws; do
    GADFIT_B200_JAC=$j timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_random_parity.py \
        tests/test_integration.py tests/test_gadf_driver.py -m gpu -q -x > $OUT/r02c_pytest_$j.log 2>&1
    tail -n 4 $OUT/r02c_pytest_$j.log
    GADFIT_B200_JAC=$j timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
        > $OUT/r02c_bench_C2_$j.json 2> $OUT/r02c_bench_C2_$j.err
    show $OUT/r02c_bench_C2_$j.json
    tail -2 $OUT/r02c_bench_C2_$j.err
done
cap() { # tag workload kernel-regex extra-args env
    GADFIT_B200_JAC=$5 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 \
        -o $OUT/r02c_prof_$1 -f python bench.py --workload $2 --steps 2 --warmup 1 --repeats 1 \
        --no-cpu-baseline --no-e2e --no-workloads $4 > $OUT/r02c_prof_$1.log 2>&1
}
cap C2_k_jacobian_sp C2 k_jacobian "" sp
cap C3_k_jacobian_sp C3 k_jacobian "" sp
cap C4_k_jacobian_sp C4 k_jacobian "--sets 64" sp
cap C5_k_jacobian_sp C5 k_jacobian "--sets 25000" sp
cap C5_k_jacobian_ws C5 k_jacobian "--sets 25000" ws
cap C5_k_chi2 C5 k_chi2 "--sets 25000" ws
cap C4_k_chi2 C4 k_chi2 "--sets 64" ws
cap C4_k_omega C4 k_omega "--sets 64" ws
ls -la $OUT | grep r02c

#!/bin/bash
# Round 2, GPU pass p: register-accumulating Jacobian sweep on config 3 (A/B against the software-pipelined one)
set -u
OUT=gpurun_out
TAG=${1:-r02p}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_random_parity.py -m gpu -q -k "not full_size_c4 and not full_size_c5" > $OUT/${TAG}_pytest_subset.log 2>&1
tail -n 4 $OUT/${TAG}_pytest_subset.log
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
      "chi2", d["final_chi2"], d.get("phase_timers"))
PY
}
for rep in 1 2; do
for j in reg sp; do
GADFIT_B200_JAC=$j timeout 600 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-workloads \
      > $OUT/${TAG}_bench_C3_${j}_$rep.json 2> $OUT/${TAG}_bench_C3.err
show $OUT/${TAG}_bench_C3_${j}_$rep.json
done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_jacobian -s 1 -c 1 \
    -o $OUT/${TAG}_prof_C3_k_jacobian_reg -f python bench.py --workload C3 --steps 2 --warmup 1 --repeats 1 \
    --no-cpu-baseline --no-e2e --no-workloads > $OUT/${TAG}_prof_C3_k_jacobian_reg.log 2>&1
tail -1 $OUT/${TAG}_prof_C3_k_jacobian_reg.log
du -sh $OUT

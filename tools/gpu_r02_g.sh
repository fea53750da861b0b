#!/bin/bash
# Round 2, seventh GPU pass: hybrid contraction (diagonal blocks on DFMA) in the warp-specialised Jacobian sweep
set -u
OUT=gpurun_out
TAG=${1:-r02g}
mkdir -p $OUT
GADFIT_PARITY_LOG=$OUT/${TAG}_parity_log.json timeout 1500 python -m pytest tests -m gpu -q \
    > $OUT/${TAG}_pytest_gpu.log 2>&1
tail -n 8 $OUT/${TAG}_pytest_gpu.log
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
      "e2e %.3g" % (d["e2e"]["value"] if d.get("e2e") else 0), "chi2", d["final_chi2"])
PY
}
timeout 900 python bench.py --steps 10 --warmup 3 --no-workloads --no-cpu-baseline > $OUT/${TAG}_bench_C2.json 2> $OUT/${TAG}_bench_C2.err
show $OUT/${TAG}_bench_C2.json
tail -3 $OUT/${TAG}_bench_C2.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_jacobian -s 1 -c 1 \
    -o $OUT/${TAG}_prof_C2_k_jacobian -f python bench.py --workload C2 --steps 2 --warmup 1 --repeats 1 \
    --no-cpu-baseline --no-e2e --no-workloads > $OUT/${TAG}_prof_C2_k_jacobian.log 2>&1
tail -2 $OUT/${TAG}_prof_C2_k_jacobian.log
du -sh $OUT

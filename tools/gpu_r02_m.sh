#!/bin/bash
# Round 2, GPU pass m: config 3 A/B -- Jacobian occupancy hint (4 CTAs/SM) and chi2 unroll 4 vs 8
set -u
OUT=gpurun_out
TAG=${1:-r02m}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fused or block_arrow_oracle or many_tiles or full_size_c3" > $OUT/${TAG}_pytest_subset.log 2>&1
tail -n 3 $OUT/${TAG}_pytest_subset.log
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
timeout 600 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-workloads \
      > $OUT/${TAG}_bench_C3_default_$rep.json 2> $OUT/${TAG}_bench_C3.err
show $OUT/${TAG}_bench_C3_default_$rep.json
GADFIT_B200_LIB=$PWD/gadfit_b200/_variants/libgadfit_b200_c3u4.so timeout 600 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-workloads \
      > $OUT/${TAG}_bench_C3_c3u4_$rep.json 2> $OUT/${TAG}_bench_C3.err
show $OUT/${TAG}_bench_C3_c3u4_$rep.json
done
du -sh $OUT

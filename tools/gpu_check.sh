#!/bin/bash
# Run on the GPU box (under gpurun): the GPU test-suite, then config 3 and
# config 2 through the fused small-problem path (and, with "ab", the
# stage-wise path as well).
#   tools/gpu_check.sh [ab]
set -u
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1
tail -n 3 $OUT/pytest_gpu.log
show() {
    python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], "ms/step %.4f" % d["ms_per_step"], "launches", d["gpu_launches"],
      {k: round(v * 1e6 / max(d["phase_calls"][k], 1), 1) for k, v in d["phase_seconds"].items()},
      "K1 us %.1f" % (d["roofline"]["seconds_per_launch"] * 1e6))
PY
}
MODES="1"
[ "${1:-}" = "ab" ] && MODES="1 0"
for w in C3 C2; do
    for f in $MODES; do
        GADFIT_B200_FUSED=$f python bench.py --workload $w --steps 10 --warmup 3 \
            --no-cpu-baseline --no-e2e > $OUT/bench_${w}_fused$f.json 2> $OUT/bench_${w}_fused$f.err
        show $OUT/bench_${w}_fused$f.json
    done
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file $OUT/launches_C3_fused.csv python bench.py --workload C3 --steps 2 --warmup 1 \
    --repeats 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1

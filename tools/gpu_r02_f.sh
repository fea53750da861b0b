#!/bin/bash
# Round 2, sixth GPU pass: single-call-site pipelined Jacobian sweep (adaptive-integral models),
# CTA tables, block-arrow inverse; full GPU suite; bench lines of all four workloads
set -u
OUT=gpurun_out
mkdir -p $OUT
for j in ws sp; do
  for w in single double; do
    GADFIT_B200_JAC=$j timeout 300 python tests/dev/sp_debug.py $w 2>&1 | tail -1
  done
done
GADFIT_PARITY_LOG=$OUT/r02f_parity_log.json timeout 1500 python -m pytest tests -m gpu -q \
    > $OUT/r02f_pytest_gpu.log 2>&1
tail -n 15 $OUT/r02f_pytest_gpu.log
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
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/r02f_bench_C2.json 2> $OUT/r02f_bench_C2.err
show $OUT/r02f_bench_C2.json
tail -3 $OUT/r02f_bench_C2.err
for j in ws sp; do
  GADFIT_B200_JAC=$j timeout 600 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-workloads \
      > $OUT/r02f_bench_C3_$j.json 2> $OUT/r02f_bench_C3_$j.err
  show $OUT/r02f_bench_C3_$j.json
  tail -3 $OUT/r02f_bench_C3_$j.err
done
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
du -sh $OUT

#!/bin/bash
# Round 2, GPU pass u: resident CTAs per SM of the omega sweep and of the Jacobian sweep of the quadrature models (A/B)
set -u
OUT=gpurun_out
mkdir -p $OUT
show() {
    python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print(sys.argv[1], "unreadable", e); sys.exit(0)
print(sys.argv[2], "ms/step %.4f" % d["ms_per_step"],
      {k: round(v * 1e6 / max(d["phase_calls"][k], 1), 1) for k, v in d["phase_seconds"].items() if k != "main"}, "chi2", d["final_chi2"])
PY
}
for v in default b3 b4; do
  for w in C4 C5; do
    if [ "$v" = "default" ]; then unset GADFIT_B200_LIB; else export GADFIT_B200_LIB=$PWD/gadfit_b200/_variants/libgadfit_b200_$v.so; fi
    timeout 600 python bench.py --workload $w --steps 6 --warmup 3 --repeats 1 --no-cpu-baseline --no-e2e --no-workloads \
        > $OUT/r02u_bench_${w}_$v.json 2> $OUT/r02u_bench.err
    show $OUT/r02u_bench_${w}_$v.json "$w $v"
  done
done

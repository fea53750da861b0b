#!/bin/bash
# Round 2, GPU pass l: shared-memory-resident k_solve_small
# captures of the config-3 kernels
set -u
OUT=gpurun_out
TAG=${1:-r02j}
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
      "e2e %.3g" % (d["e2e"]["value"] if d.get("e2e") else 0), "chi2", d["final_chi2"], d.get("phase_timers"))
for k, v in (d.get("workloads") or {}).items():
    print("   ", k, "ms/step %.4f" % v["ms_per_step"], {a: round(b * 1e6, 1) for a, b in v["seconds_per_call"].items()},
          v["phase_calls"], "launches", v["gpu_launches"], "frac %.3f (v1 %.3f)" % (v["roofline"]["frac"], v["roofline"]["frac_v1"]), v.get("phase_timers"))
PY
}
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_C2.json 2> $OUT/${TAG}_bench_C2.err
show $OUT/${TAG}_bench_C2.json
tail -3 $OUT/${TAG}_bench_C2.err
timeout 600 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-workloads \
      > $OUT/${TAG}_bench_C3.json 2> $OUT/${TAG}_bench_C3.err
show $OUT/${TAG}_bench_C3.json
cap() { # tag workload kernel-regex
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 \
        -o $OUT/${TAG}_prof_$1 -f python bench.py --workload $2 --steps 2 --warmup 1 --repeats 1 \
        --no-cpu-baseline --no-e2e --no-workloads > $OUT/${TAG}_prof_$1.log 2>&1
    tail -1 $OUT/${TAG}_prof_$1.log
}

cap C3_k_solve_small C3 k_solve_small

timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file $OUT/${TAG}_launches_C3.csv python bench.py --workload C3 --steps 10 --warmup 1 \
    --repeats 1 --no-cpu-baseline --no-e2e --no-workloads > /dev/null 2>&1
du -sh $OUT

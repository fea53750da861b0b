#!/bin/bash
# Round 2, fifth GPU pass: adaptive-integral models under both Jacobian kernels, sanitizer, pipelined upload
set -u
OUT=gpurun_out
mkdir -p $OUT
for j in ws sp; do
  for w in single double; do
    GADFIT_B200_JAC=$j timeout 300 python tests/dev/sp_debug.py $w 2>&1 | tail -1
  done
done
timeout 300 python tests/dev/sp_debug.py double oracle 2>&1 | tail -1
GADFIT_B200_JAC=sp timeout 900 compute-sanitizer --tool memcheck python tests/dev/sp_debug.py double > $OUT/r02e_memcheck_sp.log 2>&1
tail -4 $OUT/r02e_memcheck_sp.log
GADFIT_B200_JAC=sp timeout 900 compute-sanitizer --tool initcheck python tests/dev/sp_debug.py double > $OUT/r02e_initcheck_sp.log 2>&1
tail -4 $OUT/r02e_initcheck_sp.log
GADFIT_B200_JAC=sp timeout 900 compute-sanitizer --tool racecheck python tests/dev/sp_debug.py double > $OUT/r02e_racecheck_sp.log 2>&1
tail -4 $OUT/r02e_racecheck_sp.log
for j in ws sp; do
    GADFIT_B200_JAC=$j timeout 900 python -m pytest tests/test_integration.py tests/test_gadf_driver.py tests/test_model_plugin.py -m gpu -q \
        > $OUT/r02e_pytest_$j.log 2>&1
    tail -n 6 $OUT/r02e_pytest_$j.log
done
GADFIT_PARITY_LOG=$OUT/r02e_parity_log.json timeout 1500 python -m pytest tests -m gpu -q \
    --deselect tests/test_gadf_driver.py::test_device_reproduces_the_fortran_known_answers > $OUT/r02e_pytest_gpu.log 2>&1
tail -n 12 $OUT/r02e_pytest_gpu.log
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
      "e2e %.3g %.1f ms" % ((d["e2e"]["value"], d["e2e"]["seconds"] * 1e3) if d.get("e2e") else (0, 0)))
PY
}
timeout 900 python bench.py --steps 10 --warmup 3 --no-workloads --no-cpu-baseline > $OUT/r02e_bench_C2.json 2> $OUT/r02e_bench_C2.err
show $OUT/r02e_bench_C2.json
tail -3 $OUT/r02e_bench_C2.err
timeout 300 python tools/e2e_breakdown.py > $OUT/r02e_e2e_breakdown.log 2>&1
cat $OUT/r02e_e2e_breakdown.log
du -sh $OUT

#!/bin/bash
# Round 2, last GPU pass: what the driver runs at round end -- the GPU suite, smoke(), both bench arms
set -u
OUT=gpurun_out
TAG=${1:-r02last}
mkdir -p $OUT
GADFIT_PARITY_LOG=$OUT/${TAG}_parity_log.json timeout 1500 python -m pytest tests -x -q -m gpu \
    > $OUT/${TAG}_pytest_gpu.log 2>&1
tail -n 6 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > $OUT/${TAG}_bench_C2.json 2> $OUT/${TAG}_bench_C2.err
cut -c1-400 $OUT/${TAG}_bench_C2.json
tail -2 $OUT/${TAG}_bench_C2.err
for w in C3 C4 C5; do
  timeout 900 python bench.py --workload $w --no-cpu-baseline > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err
  cut -c1-250 $OUT/${TAG}_bench_$w.json
done

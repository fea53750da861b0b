#!/bin/bash
# Round 2, GPU pass t (2 GPUs): page-locked parameter gather -- multi-rank test, config 5 weak scaling line
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q > $OUT/r02t_pytest_multirank.log 2>&1
tail -n 2 $OUT/r02t_pytest_multirank.log
for w in C5 C2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --workload $w --steps 6 --warmup 3 --repeats 3 --no-cpu-baseline --no-e2e > $OUT/r02t_bench_${w}_n2.json 2> $OUT/r02t_bench_${w}_n2.err
python - $OUT/r02t_bench_${w}_n2.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["config"]["workload"][:20], "N", d["n_gpus"], "ms/step %.4f" % d["ms_per_step"], "device main %.4f s" % d["device_seconds_main_loop"], "wall/steps", d["repeats"]["seconds"], "parity", d["parity"]["ok"])
PY
done

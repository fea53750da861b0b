#!/bin/bash
# Round 2: config 5 on 8 GPUs with the final build (gpurun --gpus 8)
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
    --master-port 29561 bench.py --gpus 8 --workload C5 --steps 6 --warmup 3 --repeats 3 \
    --no-cpu-baseline --no-e2e > $OUT/bench_C5_n8.json 2> $OUT/bench_C5_n8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_C5_n8.json").read().strip().splitlines()[-1])
print("C5", d["n_gpus"], "GPUs ms/step %.3f" % d["ms_per_step"], "entries/s %.4g" % d["value"], "device main %.4f s" % d["device_seconds_main_loop"], d["repeats"]["seconds"], "parity", d["parity"]["ok"])
PY

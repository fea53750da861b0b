#!/bin/bash
# Round 2, multi-GPU pass (gpurun --gpus N): the 2-GPU parity test (N = 2) and the scaling bench line with
# its multi-rank parity record
set -u
OUT=gpurun_out
N=${1:-2}
TAG=${2:-r02}
mkdir -p $OUT
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q > $OUT/${TAG}_pytest_multirank.log 2>&1
  tail -n 3 $OUT/${TAG}_pytest_multirank.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench_C2_n$N.json 2> $OUT/${TAG}_bench_C2_n$N.err
python - $OUT/${TAG}_bench_C2_n$N.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("N", d["n_gpus"], "value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e %.4g (%.1f ms)" % (d["e2e"]["value"], d["e2e"]["seconds"] * 1e3),
      "parity ok:", d["parity"]["ok"], [(c["workload"], c["ok"], c["parameter_max_rel_diff"]) for c in d["parity"]["cases"]])
print("collectives", d["collectives"], "numa", d["config"]["host_buffers"])
PY
tail -3 $OUT/${TAG}_bench_C2_n$N.err

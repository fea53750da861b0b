#!/bin/bash
# Run on an N-GPU box (gpurun --gpus N): BASELINE configs 4 and 5 at N GPUs,
# one rank per GPU, weak scaling (every rank owns a full-size share).
#   tools/scale_configs.sh N
set -u
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
for w in C5 C4; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
        --master-port 29517 bench.py --gpus $N --workload $w --steps 6 --warmup 3 --repeats 3 \
        --no-cpu-baseline --no-e2e > $OUT/bench_${w}_n$N.json 2> $OUT/bench_${w}_n$N.err
    tail -c 1500 $OUT/bench_${w}_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w', d['n_gpus'], 'GPUs', 'ms/step %.3f'%d['ms_per_step'], 'entries/s %.4g'%d['value'], d.get('collectives'))" || tail -n 5 $OUT/bench_${w}_n$N.err
done
